set +e
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_loss_gpu.py tests/test_train_backward_gpu.py tests/test_train_forward_gpu.py tests/test_train_step_gpu.py tests/test_val_gpu.py -q -x > gpurun_out/r2_gpu_tests3.txt 2>&1
for m in 1 2 4; do Y5OBB_MSUB_MAX=$m timeout 300 python tools/time_engine.py s 16 1024 > gpurun_out/r2_te_s_msub$m.txt 2>&1; done
Y5OBB_MSUB_MAX=1 timeout 300 python tools/time_engine.py m 16 1024 > gpurun_out/r2_te_m_msub1.txt 2>&1
timeout 300 python tools/time_engine.py m 16 1024 > gpurun_out/r2_te_m_msub4.txt 2>&1
timeout 300 python tools/conv_timeline.py s 16 1024 0 1 4 16 > gpurun_out/r2_timeline4.txt 2>&1
tail -n 4 gpurun_out/r2_gpu_tests3.txt; for m in 1 2 4; do tail -n 2 gpurun_out/r2_te_s_msub$m.txt; done; tail -n 2 gpurun_out/r2_te_m_msub1.txt gpurun_out/r2_te_m_msub4.txt
