import torch, bench, time
from tests.modelgen import build_mirror
from yolov5_obb_b200.train_step import TrainStep
dev = torch.device("cuda", 0)
m = build_mirror("m", nc=15, seed=0).train().to(dev)
ts = TrainStep(m, batch_size=8, imgsz=1024)
imgs_h, tg_h = bench.train_inputs(8, 0)
imgs_d, tg_d = imgs_h.to(dev), tg_h.to(dev)
ev = []
for i in range(24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ts.step(imgs_d, tg_d)
    e1.record(); t1 = time.perf_counter()
    ev.append((e0, e1, t1 - t0))
torch.cuda.synchronize()
print(" ".join(f"{a.elapsed_time(b):.1f}/{c*1e3:.1f}" for a, b, c in ev))
