"""CPU enqueue time vs GPU time of the inference step (debug aid)."""
import time, torch, bench
from yolov5_obb_b200.general import non_max_suppression_obb
dev = torch.device("cuda", 0)
model = bench.build_model("s", dev)
x = bench.synth_batch(16, 0).to(dev)
def step():
    pred, _ = model(x)
    return non_max_suppression_obb(pred, bench.CONF, bench.IOU, multi_label=True, max_det=bench.MAX_DET, return_packed="async")
non_max_suppression_obb(model(x)[0], bench.CONF, bench.IOU, multi_label=True, max_det=bench.MAX_DET)
for _ in range(6): step()
torch.cuda.synchronize()
import cProfile, pstats
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(50): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0)/50:.3f} ms/step, total {1e3*(t2-t0)/50:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
