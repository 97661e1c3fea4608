"""GPU parity of the EXPERIMENTAL fused SGD-Nesterov + EMA step (csrc/sgd_ema.cu) against torch.optim.SGD with the
reference's parameter groups (train.py:148-162) and ModelEMA (utils/torch_utils.py:304-314) over three steps.
Skipped unless Y5OBB_EXPERIMENTAL=1: written after round 1's GPU budget was spent, not yet run on hardware."""
import copy
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("Y5OBB_EXPERIMENTAL") != "1", reason="experimental kernels: set Y5OBB_EXPERIMENTAL=1")]
DEV = "cuda:0"


def test_fused_step_matches_torch_sgd_and_ema():
    from tests.modelgen import build_mirror
    from yolov5_obb_b200.train_ops import FusedSGDEMA
    from yolov5_obb_b200.train_step import ModelEMA, param_groups
    m = build_mirror("n", nc=15, seed=0).to(DEV)
    ref = copy.deepcopy(m)
    lr, mom, wd = (0.01, 0.02, 0.1), 0.937, 5e-4
    g0, g1, g2 = param_groups(ref)
    opt = torch.optim.SGD(g0, lr=lr[0], momentum=mom, nesterov=True)
    opt.add_param_group({"params": g1, "weight_decay": wd, "lr": lr[1]})
    opt.add_param_group({"params": g2, "lr": lr[2]})
    ema_ref = ModelEMA(ref)
    ema = ModelEMA(m)
    grads = {p: torch.zeros_like(p) for p in m.parameters()}
    fused = FusedSGDEMA(m, param_groups(m), grads, ema.ema, weight_decay=wd)
    gen = torch.Generator(device=DEV).manual_seed(0)
    for step in range(3):
        for (p, g), pr in zip(grads.items(), ref.parameters()):
            g.copy_(torch.randn(p.shape, generator=gen, device=DEV))
            pr.grad = g.clone()
        with torch.no_grad():  # BatchNorm running statistics move during training: EMA follows them too
            for b, br in zip(m.buffers(), ref.buffers()):
                if b.dtype.is_floating_point:
                    b.add_(0.01)
                    br.add_(0.01)
        opt.step()
        ema_ref.update(ref)
        ema.updates += 1
        fused.step(lr, mom, ema.decay(ema.updates))
    torch.cuda.synchronize()
    for (n, a), b in zip(m.named_parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), n
    for (k, a), b in zip(ema.ema.state_dict().items(), ema_ref.ema.state_dict().values()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
