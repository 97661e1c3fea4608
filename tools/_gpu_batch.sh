set +e
timeout 280 python bench.py --no-nms-sweep --no-extra-models --no-cpu-baseline --no-eager > gpurun_out/r2_bench_final_slots.json 2> gpurun_out/r2_bench_final_slots.err
wc -l gpurun_out/r2_bench_final_slots.json; tail -3 gpurun_out/r2_bench_final_slots.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final_slots.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['single_stream']['value'], d['roofline']['frac']); t=d['train']; print(t['value'], t['ms_per_step'], t['e2e']['value'], t['tensor']['frac'], t['loss_first_last'])"
