set +e
timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_postprocess_gpu.py tests/test_pipeline_gpu.py tests/test_poly_gpu.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --no-nms-sweep --no-extra-models --no-cpu-baseline > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err
wc -l gpurun_out/r2_bench_j.json; tail -2 gpurun_out/r2_bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_j.json')); print(d['value'], d.get('step_breakdown')); print(d['train'].get('eager_torch_b200'))"
