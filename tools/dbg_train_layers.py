"""Per-layer deviation of the training forward from the fp32 oracle (debug aid)."""
import copy, sys, torch
from oracle import model_ref
from tests.modelgen import build_mirror
size, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = build_mirror(size, nc=15, seed=2).train()
ref_m = copy.deepcopy(m)
g = torch.Generator().manual_seed(8)
x = torch.rand(B, 3, H, W, generator=g)
ref, ys = model_ref.forward(ref_m, x, training=True, return_layers=True)
md = m.to("cuda:0")
with torch.no_grad():
    got = md(x.cuda())
eng = list(md._engines.values())[0]
for i, s in enumerate(eng.out_slices):
    if s is None:
        continue
    a = s.buf[..., s.c_off:s.c_off + s.C].float().cpu().permute(0, 3, 1, 2)
    b = ys[i]
    err = (a - b).abs()
    print(f"layer {i:2d} {type(md.model[i]).__name__:9s} shape {tuple(b.shape)} |ref| {b.abs().mean():.4f} mean err {err.mean():.5f} max {err.max():.4f} rel {(a-b).norm()/b.norm():.4f}")
for l in range(3):
    a, b = got[l].cpu(), ref[l]
    print("det", l, f"rel {(a-b).norm()/b.norm():.4f} mean err {(a-b).abs().mean():.5f} |ref| {b.abs().mean():.4f}")
