"""CPU tests of the host-side training logic (no kernels): the gradient-buffer write tracking of the backward plan,
the reference's optimizer parameter groups (train.py:148-156) and ModelEMA (utils/torch_utils.py:284-314)."""
import math

import pytest
import torch

from yolov5_obb_b200.conv import Slice
from yolov5_obb_b200.train_backward import _Written
from yolov5_obb_b200.train_step import ModelEMA, param_groups


def _slice(buf, off, c):
    return Slice(buf, off, c)


def test_written_ranges_decide_overwrite_or_accumulate():
    buf = torch.zeros((1, 2, 2, 64), dtype=torch.bfloat16)
    other = torch.zeros((1, 2, 2, 64), dtype=torch.bfloat16)
    w = _Written()
    assert w.contribute(_slice(buf, 0, 32)) is False            # first writer of [0, 32): overwrite
    assert w.covered(_slice(buf, 0, 32)) and not w.covered(_slice(buf, 0, 64))
    assert w.contribute(_slice(buf, 32, 32)) is False           # disjoint range: overwrite
    assert w.covered(_slice(buf, 0, 64))                        # two pieces cover the whole buffer
    assert w.contribute(_slice(buf, 16, 32)) is True            # spans both written pieces: accumulate
    assert w.contribute(_slice(other, 0, 64)) is False          # another buffer is independent
    w2 = _Written()
    w2.contribute(_slice(buf, 0, 16))
    with pytest.raises(RuntimeError):                           # partially written range: the plan refuses such a graph
        w2.contribute(_slice(buf, 8, 32))


def test_param_groups_follow_the_reference():
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, bias=False), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 4, 1))
    g0, g1, g2 = param_groups(net)
    assert [tuple(p.shape) for p in g0] == [(8,)]                       # BatchNorm weight: no decay
    assert [tuple(p.shape) for p in g1] == [(8, 3, 3, 3), (4, 8, 1, 1)]  # conv weights: decay
    assert [tuple(p.shape) for p in g2] == [(8,), (4,)]                  # biases (BN bias, conv bias)
    assert len(g0) + len(g1) + len(g2) == len(list(net.parameters()))


def test_model_ema_matches_the_reference_formula():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    ema = ModelEMA(net)
    before = {k: v.clone() for k, v in ema.ema.state_dict().items()}
    with torch.no_grad():
        for p in net.parameters():
            p.add_(1.0)
        net[1].running_mean.add_(0.5)
    ema.update(net)
    ema.update(net)
    d1, d2 = 0.9999 * (1 - math.exp(-1 / 2000)), 0.9999 * (1 - math.exp(-2 / 2000))
    msd = net.state_dict()
    for k, v in ema.ema.state_dict().items():
        if v.dtype.is_floating_point:
            want = d2 * (d1 * before[k] + (1 - d1) * msd[k]) + (1 - d2) * msd[k]
            assert torch.allclose(v, want, rtol=1e-6, atol=1e-7), k
        else:
            assert torch.equal(v, before[k])                    # integer buffers (num_batches_tracked) are not averaged
    assert ema.updates == 2 and not any(p.requires_grad for p in ema.ema.parameters())


def test_model_deepcopy_and_weight_signature():
    """copy.deepcopy(model) copies parameters but no device plans; the weight signature (what decides whether the inference
    plan must be rebuilt) moves with in-place updates of parameters and of num_batches_tracked, and only then."""
    import copy
    from tests.modelgen import build_mirror
    m = build_mirror("n", nc=15, seed=0)
    m._engines["plan"] = object()            # stands for a device plan (holds ctypes handles: not copyable)
    m._last_train_engine = object()
    s0 = m._weights_signature()
    assert m._weights_signature() == s0
    c = copy.deepcopy(m)
    assert c._engines == {} and not hasattr(c, "_last_train_engine")
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    sc = c._weights_signature()
    with torch.no_grad():
        next(c.parameters()).add_(1.0)       # an optimizer / EMA step on the copy
    assert c._weights_signature() != sc and m._weights_signature() == s0
    bns = [b for b in m.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    torch._foreach_add_([b.num_batches_tracked for b in bns], 1)   # what a training forward does
    assert m._weights_signature() != s0
    m.invalidate()
    assert m._engines == {}
