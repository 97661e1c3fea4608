set +e
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/mma_rate tools/micro/mma_rate.cu > /dev/null 2>&1
timeout 300 ./tools/micro/mma_rate > gpurun_out/r2_mma_rate2.txt 2>&1
Y5OBB_CONV_FLAGS=1280 timeout 300 python tools/conv_timeline.py s 16 1024 0 2 3 16 > gpurun_out/r2_timeline_mmaonly.txt 2>&1
Y5OBB_CONV_FLAGS=768 timeout 300 python tools/conv_timeline.py s 16 1024 0 2 3 16 > gpurun_out/r2_timeline_epionly.txt 2>&1
timeout 300 python tools/conv_timeline.py s 16 1024 16 22 > gpurun_out/r2_timeline3.txt 2>&1
cat gpurun_out/r2_mma_rate2.txt
