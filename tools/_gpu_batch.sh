set +e
(timeout 120 python -m pytest tests/test_train_step_gpu.py -q -k "run_over_host" 2>&1 | tail -40) > gpurun_out/r2_trainrun_test.txt
timeout 200 python bench.py --no-nms-sweep --no-extra-models --no-cpu-baseline --no-eager --no-parity-gate --steps 30 > gpurun_out/r2_bench_trainrun.json 2> gpurun_out/r2_bench_trainrun.err
tail -5 gpurun_out/r2_trainrun_test.txt
wc -l gpurun_out/r2_bench_trainrun.json; tail -3 gpurun_out/r2_bench_trainrun.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_trainrun.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['single_stream']['value'], d['roofline']['frac']); t=d['train']; print(t['value'], t['ms_per_step'], t['e2e']['value'], t['e2e']['ms_per_step'], t['loss_first_last'])"
