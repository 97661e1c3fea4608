import torch
from tests.modelgen import build_mirror
DEV = "cuda:0"
m = build_mirror("n", nc=15, seed=4).train().to(DEV)
g = torch.Generator().manual_seed(5)
x = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
outs = m(x)
G = [torch.randn(o.shape, generator=g).to(DEV) * 0.05 for o in outs]
sum((o * gg).sum() for o, gg in zip(outs, G)).backward()
g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
outs = m(x)
sum((o * gg).sum() for o, gg in zip(outs, G)).backward()
torch.cuda.synchronize()
acc = {n: p.grad.clone() for n, p in m.named_parameters()}
m.zero_grad()
outs = m(x)
sum((o * gg).sum() for o, gg in zip(outs, G)).backward()
g3 = {n: p.grad.clone() for n, p in m.named_parameters()}
bad = 0
for n in g1:
    r_acc = ((acc[n] - 2 * g1[n]).norm() / g1[n].norm()).item()
    r_3 = ((g3[n] - g1[n]).norm() / g1[n].norm()).item()
    if r_acc > 1e-3 or r_3 > 1e-3:
        bad += 1
        if bad < 12:
            print(f"{n}: acc-vs-2g1 {r_acc:.4f}  g3-vs-g1 {r_3:.4f}")
print("bad", bad, "of", len(g1))
w = "model.0.conv.weight"
for dy in range(2):
    for dx in range(2):
        a, b = acc[w][:, :, dy::2, dx::2], 2 * g1[w][:, :, dy::2, dx::2]
        print("stem phase", dy, dx, ((a - b).norm() / b.norm()).item(), ((g3[w][:, :, dy::2, dx::2] - b / 2).norm() / b.norm()).item())
