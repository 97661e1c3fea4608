"""GPU parity of the training-mode forward (batch-statistic BatchNorm, raw Detect outputs) against the fp32
oracle (oracle/model_ref.forward(training=True), itself pinned to the REFERENCE model in eval mode and sharing
its code path).  Tolerance as in test_engine_gpu (bf16 activations): 3e-2 rel + 3e-2 abs, mean error < 6e-3;
running statistics must match torch.nn.BatchNorm2d's update within 2e-2 relative."""
import copy

import pytest
import torch

from oracle import model_ref
from tests.modelgen import build_mirror

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("size,B,H,W", [("n", 2, 128, 160), ("s", 2, 64, 64)])
def test_train_forward_matches_oracle(size, B, H, W):
    m = build_mirror(size, nc=15, seed=2).train()
    ref_m = copy.deepcopy(m)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(B, 3, H, W, generator=g)
    ref = model_ref.forward(ref_m, x, training=True)   # updates ref_m's running stats
    md = m.to(DEV)
    with torch.no_grad():
        got = md(x.to(DEV))
    torch.cuda.synchronize()
    assert isinstance(got, list) and len(got) == 3
    for l, (a, b) in enumerate(zip(got, ref)):
        assert a.shape == b.shape and a.dtype == torch.float32
        err = (a.cpu() - b).abs()
        tol = b.abs() * 3e-2 + 3e-2
        frac = (err > tol).float().mean().item()
        print(f"level {l}: max err {err.max().item():.4g} mean {err.mean().item():.4g} out-of-tol {frac:.2e}")
        assert frac < 2e-3 and err.mean().item() < 6e-3
    bns = [(n, mod) for n, mod in md.named_modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    refs = dict(ref_m.named_modules())
    worst = 0.0
    for n, mod in bns:
        r = refs[n]
        assert int(mod.num_batches_tracked) == int(r.num_batches_tracked) == 1
        dm = (mod.running_mean.cpu() - r.running_mean).abs().max().item()
        dv = ((mod.running_var.cpu() - r.running_var).abs() / r.running_var.abs().clamp_min(1e-3)).max().item()
        worst = max(worst, dm, dv)
    print("worst running-stat deviation", worst)
    assert worst < 2e-2
