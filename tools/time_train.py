"""Phase timing of one training step (forward / loss / backward / optimiser+EMA) with CUDA events (debug aid).
usage: python tools/time_train.py [size=m] [B=8] [imgsz=1024] [steps=5]"""
import sys, time, torch
from tests.modelgen import build_mirror
from tests.tilegen import synth_tiles
from tests.lossgen import synth_targets
from yolov5_obb_b200.train_step import TrainStep

size = sys.argv[1] if len(sys.argv) > 1 else "m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
warm = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = "cuda:0"
m = build_mirror(size, nc=15, seed=0).train().to(dev)
ts = TrainStep(m, batch_size=64)
imgs = synth_tiles(B, S, seed=1).to(dev)
tg = torch.from_numpy(synth_targets(B, 24 * B, S, nc=15, seed=2)).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(warm + steps):
    e = [ev() for _ in range(5)]
    t0 = time.perf_counter()
    e[0].record()
    pred = m(imgs)
    e[1].record()
    loss, items = ts.compute_loss(pred, tg)
    e[2].record()
    loss.backward()
    e[3].record()
    ts.optimizer.step(); ts.optimizer.zero_grad(set_to_none=True); ts.ema.update(m)
    e[4].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if it >= warm:
        d = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
        print(f"step {it}: fwd {d[0]:.2f} loss {d[1]:.2f} bwd {d[2]:.2f} opt+ema {d[3]:.2f} total {sum(d):.2f} ms (wall {wall:.2f})  "
              f"{B / sum(d) * 1e3:.1f} img/s  loss {loss.item():.4f}", flush=True)
eng = m._last_train_engine
print("fwd flops %.3g  bwd flops %.3g  mem %.2f GB" % (eng.flops, eng._bwd.flops, torch.cuda.max_memory_allocated() / 2**30))
