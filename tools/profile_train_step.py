"""One steady-state training step under the CUDA profiler API (for `ncu --profile-from-start off`): builds the yolov5m b8 1024^2
plan, runs warm-up steps (plan building and its one-off allocations / fills stay outside), then exactly `steps` TrainStep.step calls
between cudaProfilerStart / Stop.  Y5OBB_NO_GRAPH=1 makes the launches individually visible.
usage: python tools/profile_train_step.py [size=m] [B=8] [imgsz=1024] [steps=1]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.lossgen import synth_targets
from tests.modelgen import build_mirror
from tests.tilegen import synth_tiles
from yolov5_obb_b200.train_step import TrainStep

size = sys.argv[1] if len(sys.argv) > 1 else "m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = "cuda:0"
m = build_mirror(size, nc=15, seed=0).train().to(dev)
ts = TrainStep(m, batch_size=64, imgsz=S, warmup_iters=1000)
imgs = synth_tiles(B, S, seed=1).to(dev)
tg = torch.from_numpy(synth_targets(B, 24 * B, S, nc=15, seed=2)).to(dev)
for _ in range(4):
    ts.step(imgs, tg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    loss, _ = ts.step(imgs, tg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
