set +e
Y5OBB_DUAL=2 timeout 300 python tools/time_engine.py s 16 1024 > gpurun_out/r2_te_s_dual2.txt 2>&1
Y5OBB_DUAL=2 timeout 300 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2_gpu_tests10.txt 2>&1
tail -n 2 gpurun_out/r2_te_s_dual2.txt; grep -c "grid=296" gpurun_out/r2_te_s_dual2.txt; tail -n 2 gpurun_out/r2_gpu_tests10.txt
