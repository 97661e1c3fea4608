"""CPU test: the non_max_suppression_obb restatement against outputs of the REFERENCE function
(tests/golden/postprocess_golden.npz, made by make_postprocess_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle.postprocess import non_max_suppression_obb
from tests.predgen import synth_pred
from tests.golden_cfgs import PP_CFGS

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("name", sorted(PP_CFGS))
def test_postprocess_oracle_matches_reference_output(name):
    G = np.load(ROOT / "tests" / "golden" / "postprocess_golden.npz")
    pred = torch.from_numpy(synth_pred(2, int(G[f"{name}/anchors"]), 15, int(G[f"{name}/seed"])))
    for mode in (0, 1):
        res = non_max_suppression_obb(pred, nms_mode=mode, **PP_CFGS[name])
        for b, r in enumerate(res):
            assert np.array_equal(r.numpy(), G[f"{name}/{b}"]), (name, mode, b)


def test_postprocess_oracle_over_max_nms_order():
    """> 30 000 candidates per image with degenerate boxes ranked inside the top 30 000: the reference clamps by score
    first (general.py:845-846) and drops min(w,h) < 0.001 afterwards (nms_rotated_wrapper.py:32-39)."""
    from tests.golden_cfgs import PP_OVERMAX
    G = np.load(ROOT / "tests" / "golden" / "postprocess_golden.npz")
    pred = torch.from_numpy(synth_pred(seed=int(G["overmax/seed"]), **PP_OVERMAX["pred"]))
    res = non_max_suppression_obb(pred, nms_mode=1, **PP_OVERMAX["kw"])
    for b, r in enumerate(res):
        assert np.array_equal(r.numpy(), G[f"overmax/{b}"]), b
    # the other order (drop first, clamp second) gives a different answer on this input: the case discriminates
    tiny = pred[..., 2:4].min(-1)[0] < 0.001
    other = pred.clone()
    other[..., 4] = torch.where(tiny, torch.zeros(()), other[..., 4])
    res2 = non_max_suppression_obb(other, nms_mode=1, **PP_OVERMAX["kw"])
    assert any(a.shape != b.shape for a, b in zip(res, res2))
