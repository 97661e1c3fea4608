set +e
Y5OBB_TE_RECORDS=1 timeout 300 python tools/conv_timeline.py s 16 1024 0 1 4 5 9 > gpurun_out/r2_tl_epi.txt 2>&1
grep -c . gpurun_out/r2_tl_epi.txt
