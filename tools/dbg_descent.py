import torch, sys
from tests.lossgen import synth_targets
from tests.modelgen import build_mirror
from tests.tilegen import synth_tiles
from yolov5_obb_b200.train_step import TrainStep, HYP_FINETUNE_DOTA
DEV = "cuda:0"
B, S = 4, 128
imgs = synth_tiles(B, S, seed=3).to(DEV)
tg = torch.from_numpy(synth_targets(B, 40, S, nc=15, seed=3)).to(DEV)
for lr in (1e-2, 1e-3, 1e-4, 1e-5):
    m = build_mirror("n", nc=15, seed=1).train().to(DEV)
    hyp = dict(HYP_FINETUNE_DOTA); hyp["lr0"] = lr
    ts = TrainStep(m, hyp=hyp, batch_size=64)
    ls = []
    for it in range(16):
        l, items = ts.step(imgs, tg)
        ls.append(l.item())
    print(lr, " ".join(f"{v:.3f}" for v in ls), flush=True)
    gn = 0
