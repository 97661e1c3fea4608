"""Per-op CUDA-event timing of one inference plan (dev tool): python tools/time_engine.py s 16 1024"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.modelgen import build_mirror
from yolov5_obb_b200 import _lib
from yolov5_obb_b200.engine import InferenceEngine

size, B, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
import os
if os.environ.get("Y5OBB_TE_CALIBRATED") == "1":   # the benchmark's 'alive' model (tests/modelgen.calibrated_bench_model)
    import bench
    m = bench.build_model(size).cuda()
    x = bench.synth_batch(B, 0).cuda()
else:
    m = build_mirror(size, nc=15, seed=0).cuda()
    x = None
eng = InferenceEngine(m, B, S, S, torch.device("cuda:0"), conv_flags=int(os.environ.get("Y5OBB_CONV_FLAGS", "0")),
                      compact_detect=os.environ.get("Y5OBB_TE_RECORDS") == "1")   # 1: the bench's plan (Detect rows as compact records)
if x is None:
    x = torch.rand(B, 3, S, S, device="cuda")
for _ in range(3):
    eng.forward(x)
torch.cuda.synchronize()
st = _lib.stream_ptr(eng.device)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.ops) + 1)]
evs[0].record()
for i, op in enumerate(eng.ops):
    op(st)
    evs[i + 1].record()
torch.cuda.synchronize()
tot = 0.0
ci = 0
for i in range(len(eng.ops)):
    ms = evs[i].elapsed_time(evs[i + 1])
    tot += ms
    info = ""
    if ci < len(eng.convs) and eng.ops[i].__defaults__ and eng.ops[i].__defaults__[0] is eng.convs[ci]._h:
        inf = eng.convs[ci].info()
        info = f"conv BN={inf['block_n']} BK={inf['block_k']} st={inf['stages']} grid={inf['grid']} " \
               f"{inf['flops'] / ms / 1e9:8.1f} TF/s {inf['hbm_bytes'] / ms / 1e6:8.1f} GB/s"
        ci += 1
    print(f"{i:3d} {ms * 1000:9.1f} us  {info}")
print(f"sum of ops {tot:.3f} ms; flops {eng.flops / 1e9:.1f} G -> {eng.flops / tot / 1e9:.1f} TF/s; "
      f"alg bytes {eng.hbm_bytes / 1e6:.0f} MB -> {eng.hbm_bytes / tot / 1e6:.0f} GB/s")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    eng.forward(x)
e.record()
torch.cuda.synchronize()
print(f"forward {s.elapsed_time(e) / 10:.3f} ms -> {B / (s.elapsed_time(e) / 10) * 1000:.0f} img/s")
