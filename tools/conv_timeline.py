"""In-kernel timeline of conv_tc_kernel (dev tool): CTA 0 stamps clock64() in its producer / MMA / epilogue roles for its
first 32 tiles (y5obb_conv_debug_timestamps).  python tools/conv_timeline.py s 16 1024 [layer indices...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.modelgen import build_mirror
from yolov5_obb_b200 import _lib
from yolov5_obb_b200.engine import InferenceEngine

size, B, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
which = [int(a) for a in sys.argv[4:]] or None
m = build_mirror(size, nc=15, seed=0).cuda()
import os
eng = InferenceEngine(m, B, S, S, torch.device("cuda:0"), conv_flags=int(os.environ.get("Y5OBB_CONV_FLAGS", "0")),
                      compact_detect=os.environ.get("Y5OBB_TE_RECORDS") == "1")
x = torch.rand(B, 3, S, S, device="cuda")
for _ in range(2):
    eng.forward(x)
torch.cuda.synchronize()
L = _lib.lib()
st = _lib.stream_ptr(eng.device)
buf = torch.zeros(3 * 32 * 8, dtype=torch.int64, device="cuda")
for ci, cv in enumerate(eng.convs):
    if which is not None and ci not in which:
        continue
    inf = cv.info()
    buf.zero_()
    L.y5obb_conv_debug_timestamps(cv._h, buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cv.run(st)
    e1.record()
    torch.cuda.synchronize()
    L.y5obb_conv_debug_timestamps(cv._h, None)
    t = buf.cpu().view(3, 32, 8)
    t0 = int(t[..., :7][t[..., :7] > 0].min())
    order = sorted((int(t[1, k, 7]), k) for k in range(32) if int(t[1, k, 7]))
    print(f"--- conv {ci}: BN={inf['block_n']} BK={inf['block_k']} stages={inf['stages']} grid={inf['grid']} launch {e0.elapsed_time(e1) * 1e3:.1f} us "
          f"(the CTA's LAST 32 tiles; cycles since the earliest surviving stamp; P=producer [wait-empty, got, loads issued], M=mma [wait-tmem, got, "
          f"first full, issued, committed], E=epilogue [wait-full, got, done, arrived | first chunk: before tcgen05.ld, after wait::ld, staged]; E rows: the epilogue group that owns the tile)")
    for ordinal, it in order:
        f = lambda r, n: " ".join(f"{int(t[r, it, k]) - t0:7d}" if int(t[r, it, k]) else "      -" for k in range(n))
        print(f"tile {ordinal - 1:3d}  P {f(0, 3)} | M {f(1, 5)} | E {f(2, 4)} | {" ".join(f"{int(t[2, it, k]) - t0:7d}" if int(t[2, it, k]) else "      -" for k in (4, 5, 6))}")
