set +e
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_final.txt 2>&1
tail -n 3 gpurun_out/r2_gpu_tests_final.txt
timeout 1200 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
cut -c1-300 gpurun_out/r2_bench_final2.json; tail -2 gpurun_out/r2_bench_final2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke2.txt 2>&1; tail -2 gpurun_out/r2_smoke2.txt
