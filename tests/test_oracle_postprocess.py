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
