set +e
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -4
Y5OBB_NO_GRAPH=1 Y5OBB_TE_CALIBRATED=1 Y5OBB_TE_RECORDS=1 timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 159 -c 53 -o /tmp/r2_conv_full python tools/time_engine.py s 16 1024 > gpurun_out/r2_ncu_c.log 2>&1
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_ncu_raw.csv 2>/dev/null
Y5OBB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 260 --csv --log-file gpurun_out/r2_launches_infer.csv python bench.py --steps 3 --warmup 3 --no-train --no-eager --no-nms-sweep --no-cpu-baseline --no-parity-gate --no-extra-models > gpurun_out/r2_ncu_b.log 2>&1
wc -l gpurun_out/r2_conv_ncu_raw.csv gpurun_out/r2_launches_infer.csv
