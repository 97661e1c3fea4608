set +e
timeout 900 python -m pytest tests/test_val_gpu.py tests/test_pipeline_gpu.py tests/test_postprocess_gpu.py tests/test_engine_gpu.py tests/test_train_ops_gpu.py tests/test_train_backward_gpu.py -q > gpurun_out/r2_gpu_tests6.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-eager --no-nms-sweep --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
Y5OBB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 260 --csv --log-file gpurun_out/r2_launches_infer.csv python bench.py --steps 3 --warmup 3 --no-train --no-eager --no-nms-sweep --no-cpu-baseline --no-parity-gate > gpurun_out/r2_ncu_b.log 2>&1
tail -n 3 gpurun_out/r2_gpu_tests6.txt; head -c 300 gpurun_out/r2_bench_d.json; tail -n 2 gpurun_out/r2_bench_d.err
