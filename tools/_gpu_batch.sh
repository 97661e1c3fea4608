set +e
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_nms_gpu.py tests/test_postprocess_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-train --no-eager --no-nms-sweep --no-extra-models --no-cpu-baseline > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err
cut -c1-160 gpurun_out/r2_bench_i.json; tail -2 gpurun_out/r2_bench_i.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_i.json')); print(d.get('step_breakdown'))"
