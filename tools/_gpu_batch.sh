set +e
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_train_forward_gpu.py tests/test_train_backward_gpu.py tests/test_train_step_gpu.py -q -x 2>&1 | tail -3
Y5OBB_TE_CALIBRATED=1 Y5OBB_TE_RECORDS=1 timeout 300 python tools/time_engine.py s 16 1024 > gpurun_out/r2_te_epi6.txt 2>&1
tail -n 2 gpurun_out/r2_te_epi6.txt
timeout 600 python bench.py --no-train --no-eager --no-nms-sweep --no-extra-models --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
cut -c1-200 gpurun_out/r2_bench_g.json; tail -2 gpurun_out/r2_bench_g.err
