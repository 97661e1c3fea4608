set +e
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_slots.txt
cat gpurun_out/r2_gpu_tests_slots.txt
