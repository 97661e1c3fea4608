set +e
timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_postprocess_gpu.py -q -x > gpurun_out/r2_gpu_tests11.txt 2>&1
timeout 600 python - > gpurun_out/r2_nms_sweep.json 2> gpurun_out/r2_nms_sweep.err <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
torch.zeros(1, device='cuda')
print(json.dumps(bench.nms_sweep(torch.device('cuda:0'))))
PY
Y5OBB_NO_GRAPH=1 Y5OBB_TE_CALIBRATED=1 Y5OBB_TE_RECORDS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 159 -c 53 -o /tmp/r2_conv_full python tools/time_engine.py s 16 1024 > gpurun_out/r2_ncu_c.log 2>&1
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_ncu_raw.csv 2>/dev/null
Y5OBB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 260 --csv --log-file gpurun_out/r2_launches_infer.csv python bench.py --steps 3 --warmup 3 --no-train --no-eager --no-nms-sweep --no-cpu-baseline --no-parity-gate --no-extra-models > gpurun_out/r2_ncu_b.log 2>&1
tail -n 3 gpurun_out/r2_gpu_tests11.txt; tail -n 2 gpurun_out/r2_nms_sweep.err; python - <<'PY'
import json
try:
    n=json.load(open('gpurun_out/r2_nms_sweep.json')); print(n['value'], n['all_keep_lists_equal_reference_k1'])
    for x in n['rows']:
        if x['n']>=10000: print(x['layout'],x['n'],round(x['ms'],3),x['stage_ms'],x['keep_list_equals_reference_k1'])
except Exception as e: print('sweep failed',e)
PY
wc -l gpurun_out/r2_conv_ncu_raw.csv gpurun_out/r2_launches_infer.csv
