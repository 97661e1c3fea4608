"""CPU test: the loss restatement (oracle/loss_ref.py, autograd for gradients) against outputs of the
REFERENCE ComputeLoss (tests/golden/loss_golden.npz, made by make_loss_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import loss_ref
from tests.lossgen import synth_preds
from tests.losscases import CASES, hyp_from_golden, ANCHORS_GRID, STRIDES

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("name", sorted(CASES))
def test_loss_oracle_matches_reference(name):
    G = np.load(ROOT / "tests" / "golden" / "loss_golden.npz")
    c = CASES[name]
    p = [x.requires_grad_(True) for x in synth_preds(c["B"], c["imgsz"], seed=c["seed"])]
    tg = torch.from_numpy(G[f"{name}/targets"])
    hyp = hyp_from_golden(G, name)
    loss, items = loss_ref.compute_loss(p, tg, ANCHORS_GRID, STRIDES, hyp, nc=15)
    np.testing.assert_allclose(loss.detach().numpy(), G[f"{name}/loss"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(items.numpy(), G[f"{name}/items"], rtol=1e-6, atol=1e-6)
    (loss * 3.0).backward()
    for i, x in enumerate(p):
        g = x.grad.reshape(-1, 200)
        np.testing.assert_allclose(g[:, 4].numpy(), G[f"{name}/gobj{i}"], rtol=1e-5, atol=1e-8)
        rows = G[f"{name}/grows{i}"]
        np.testing.assert_allclose(g[rows].numpy(), G[f"{name}/gvals{i}"], rtol=1e-5, atol=1e-8)
