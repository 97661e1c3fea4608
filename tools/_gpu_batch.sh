set +e
export PYTHONPATH=.
(timeout 400 python -m pytest tests/test_train_backward_gpu.py tests/test_train_step_gpu.py tests/test_sgd_ema_gpu.py -q 2>&1 | tail -8) > gpurun_out/r2_wgrad_stream.txt
echo "--- Y5OBB_WGRAD_STREAM=0" >> gpurun_out/r2_wgrad_stream.txt
(Y5OBB_WGRAD_STREAM=0 timeout 200 python tools/time_train.py m 8 1024 4 2>&1 | head -6) >> gpurun_out/r2_wgrad_stream.txt
echo "--- Y5OBB_WGRAD_STREAM=1" >> gpurun_out/r2_wgrad_stream.txt
(Y5OBB_WGRAD_STREAM=1 timeout 200 python tools/time_train.py m 8 1024 4 2>&1 | head -6) >> gpurun_out/r2_wgrad_stream.txt
cat gpurun_out/r2_wgrad_stream.txt
