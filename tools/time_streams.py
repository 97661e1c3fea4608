"""Dev tool: does keeping several batches in flight on separate streams (one plan + one post-process graph per slot) raise the
throughput of the inference step (pre-process + forward + NMS, yolov5s b16 1024^2, the bench workload)?
    python tools/time_streams.py [steps]
Prints ms per step for 1, 2 and 3 slots and checks that every slot produces the single-stream result bit for bit."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from yolov5_obb_b200.general import non_max_suppression_obb

K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
size = sys.argv[2] if len(sys.argv) > 2 else "s"
B = 16
dev = torch.device("cuda", 0)
model = bench.build_model(size).to(dev)
x = bench.synth_batch(B, seed=0).to(dev)
kw = dict(conf_thres=bench.CONF, iou_thres=bench.IOU, multi_label=True, max_det=bench.MAX_DET)

# sizes the candidate capacity (sticky hint) before any async call
ref = non_max_suppression_obb(model.detect_records(x), **kw)


def step(slot):
    rec = model.detect_records(x, slot=slot)
    return non_max_suppression_obb(rec, return_packed="async", **kw)


def run(n_slots, steps):
    main = torch.cuda.current_stream(dev)
    streams = [main] if n_slots == 1 else [torch.cuda.Stream(dev) for _ in range(n_slots)]
    # warm every slot alone (plan build, the two eager calls before each graph capture)
    outs = [None] * n_slots
    for s in range(n_slots):
        with torch.cuda.stream(streams[s]):
            for _ in range(4):
                outs[s] = step(s)
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams:
        if st is not main:
            st.wait_event(e0)
    for i in range(steps):
        s = i % n_slots
        with torch.cuda.stream(streams[s]):
            outs[s] = step(s)
    for st in streams:
        if st is not main:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, outs


base = None
for n in (1, 2, 3, 2, 1):
    ms, outs = run(n, K)
    same = True
    for o in outs:
        c = o[1].tolist()
        dets = [o[0][b, :c[b]] for b in range(B)]
        same = same and all(torch.equal(a, b) for a, b in zip(dets, ref))
    print(f"slots {n}: {ms:.4f} ms per step -> {B / ms * 1e3:.0f} img/s; results equal the blocking call's: {same}", flush=True)
