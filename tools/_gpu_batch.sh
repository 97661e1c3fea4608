set +e
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_postprocess_gpu.py tests/test_val_gpu.py tests/test_train_forward_gpu.py -q > gpurun_out/r2_gpu_tests4.txt 2>&1
timeout 300 python tools/time_engine.py s 16 1024 > gpurun_out/r2_te_s_b.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-nms-sweep > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -n 4 gpurun_out/r2_gpu_tests4.txt; tail -n 2 gpurun_out/r2_te_s_b.txt; head -c 400 gpurun_out/r2_bench_b.json; tail -n 3 gpurun_out/r2_bench_b.err
