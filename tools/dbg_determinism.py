import torch
from tests.modelgen import build_mirror
DEV = "cuda:0"
m = build_mirror("n", nc=15, seed=4).train().to(DEV)
g = torch.Generator().manual_seed(5)
x = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
res = []
for it in range(3):
    m.zero_grad()
    outs = m(x)
    if it == 0:
        G = [torch.randn(o.shape, generator=g).to(DEV) * 0.05 for o in outs]
    o_copy = [o.clone() for o in outs]
    sum((o * gg).sum() for o, gg in zip(outs, G)).backward()
    torch.cuda.synchronize()
    res.append((o_copy, {n: p.grad.clone() for n, p in m.named_parameters()}))
for it in (1, 2):
    print("run", it, "outs rel", [((a - b).norm() / b.norm()).item() for a, b in zip(res[it][0], res[0][0])])
    rels = sorted((((res[it][1][n] - res[0][1][n]).norm() / res[0][1][n].norm()).item(), n) for n in res[0][1])
    print("  grads: median", rels[len(rels) // 2], "max", rels[-1])
