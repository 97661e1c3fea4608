"""CPU: the oracle's TRAIN-mode forward and its autograd gradients (oracle/model_ref.forward_with_grad, the checker of
the device training path) against the REFERENCE model itself run in train mode with the same seed
(tests/golden/train_golden.npz, make_train_golden.py): outputs within 1e-4, BatchNorm running statistics within 1e-5,
every parameter gradient's norm / probe projection within 1e-3 relative and its leading elements within 1e-3 of the
gradient's scale (fp32 summation order differs between the two graphs)."""
from pathlib import Path

import numpy as np
import torch

from oracle import model_ref
from tests.modelgen import build_mirror

ROOT = Path(__file__).resolve().parents[1]


def test_oracle_train_mode_matches_reference_model():
    G = np.load(ROOT / "tests" / "golden" / "train_golden.npz")
    m = build_mirror("n", nc=15, seed=0).train()
    x = torch.from_numpy(G["x"])
    with torch.enable_grad():
        outs = model_ref.forward_with_grad(m, x, training=True)
        loss = sum((o * torch.from_numpy(G[f"G{l}"])).sum() for l, o in enumerate(outs))
        loss.backward()
    for l, o in enumerate(outs):
        ref = G[f"out{l}"]
        assert o.shape == ref.shape
        assert np.abs(o.detach().numpy() - ref).max() < 1e-4
    names = [str(n) for n in G["names"]]
    params = dict(m.named_parameters())
    assert names == list(params)                       # same parameter names and order as the reference modules
    worst = 0.0
    for i, n in enumerate(names):
        ref = G[f"grad/{n}"]
        g = params[n].grad.double().flatten()
        probe = torch.randn(g.numel(), generator=torch.Generator().manual_seed(1000 + i), dtype=torch.float64)
        scale = max(ref[0], 1e-12)
        assert abs(g.norm().item() - ref[0]) < 1e-3 * scale, n
        assert abs((g * probe).sum().item() - ref[1]) < 1e-3 * scale * np.sqrt(g.numel()), n
        lead = g[:64].numpy()
        worst = max(worst, np.abs(lead - ref[2:2 + lead.size]).max() / (scale / np.sqrt(g.numel()) + 1e-12))
    assert worst < 5e-2, worst                          # element error relative to the RMS element
    bufs = dict(m.named_buffers())
    for k in G.files:
        if k.startswith("buf/"):
            assert np.abs(bufs[k[4:]].numpy() - G[k]).max() < 1e-5, k
