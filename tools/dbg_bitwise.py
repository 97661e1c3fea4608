"""Which backward buffers are not bitwise reproducible between identical runs? (debug aid)"""
import sys, torch
from tests.modelgen import build_mirror
DEV = "cuda:0"
size, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = build_mirror(size, nc=15, seed=4).train().to(DEV)
g = torch.Generator().manual_seed(5)
x = torch.rand(B, 3, H, W, generator=g).to(DEV)
names = {id(mm): n for n, mm in m.named_modules()}
snaps = []
for it in range(4):
    m.zero_grad()
    outs = m(x)
    if it == 0:
        G = [torch.randn(o.shape, generator=g).to(DEV) * 0.05 for o in outs]
    sum((o * gg).sum() for o, gg in zip(outs, G)).backward()
    torch.cuda.synchronize()
    eng = m._last_train_engine
    plan = eng._bwd
    snap = {}
    for part in plan.conv_parts:
        lay = part["lay"]
        snap[("dz", names[id(lay.mod)])] = part["dz"].clone()
        snap[("z", names[id(lay.mod)])] = lay.z.buf.clone()
        snap[("y", names[id(lay.mod)])] = lay.y.buf.clone()
        gy = plan.gbuf[lay.y.buf.data_ptr()]
        snap[("gybuf", names[id(lay.mod)])] = gy.clone()
    for part in plan.det_parts:
        snap[("dzd", part["l"])] = part["dzd"].clone()
    snap[("flat", "")] = plan.flat.clone()
    snaps.append(snap)
for it in range(1, 4):
    diff = [(k, (snaps[it][k].float() - snaps[0][k].float()).abs().max().item()) for k in snaps[0] if not torch.equal(snaps[it][k], snaps[0][k])]
    print("run", it, "differs in", len(diff), "buffers:", [(k, f"{v:.3g}") for k, v in diff[:12]])
