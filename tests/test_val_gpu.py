"""GPU parity of the fused post-NMS geometry + validation matching kernel (csrc/val_match.cu) against outputs
of the REFERENCE functions chained as val.py:226-250 does (tests/golden/valmatch_golden.npz): `correct` matrices equal,
native-space polygons and HBB boxes bit-exact (every step is a separately rounded fp32 op, like the ATen chain)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.valgen import synth_val_batch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
DEV = "cuda:0"


@pytest.mark.parametrize("seed", [0, 1])
def test_val_matching_equals_reference(seed):
    from yolov5_obb_b200.val import match_batch
    G = np.load(ROOT / "tests" / "golden" / "valmatch_golden.npz")
    dets, counts, targets, shapes = synth_val_batch(seed)
    iouv = torch.linspace(0.5, 0.95, 10)
    correct, polyn, hbbn = match_batch(torch.from_numpy(dets).to(DEV), torch.from_numpy(counts).to(DEV),
                                       torch.from_numpy(targets).to(DEV), shapes, iouv.to(DEV))
    assert correct.dtype == torch.bool and correct.shape == (3, 300, 10)
    for si in range(3):
        n = int(counts[si])
        want = G[f"{seed}/{si}/correct"]
        got = correct[si, :n].cpu().numpy()
        assert np.array_equal(got, want), (si, int((got != want).sum()))
        assert not correct[si, n:].any()
        assert np.array_equal(polyn[si, :n].cpu().numpy().view(np.uint32), G[f"{seed}/{si}/polyn"].view(np.uint32)), si
        assert np.array_equal(hbbn[si, :n].cpu().numpy().view(np.uint32), G[f"{seed}/{si}/hbbn"].view(np.uint32)), si


def test_val_matching_on_the_nms_output():
    """The packed (dets, counts) pair of non_max_suppression_obb(..., return_packed=True) feeds the kernel directly."""
    from tests.predgen import synth_pred
    from yolov5_obb_b200.general import non_max_suppression_obb
    from yolov5_obb_b200.val import match_batch
    pred = torch.from_numpy(synth_pred(3, 3000, 15, 3)).to(DEV)
    out, rows = non_max_suppression_obb(pred, 0.25, 0.45, multi_label=True, max_det=300, return_packed=True)
    _, _, targets, shapes = synth_val_batch(0)
    correct, polyn, hbbn = match_batch(out, torch.tensor(rows, device=DEV), torch.from_numpy(targets).to(DEV), shapes,
                                       torch.linspace(0.5, 0.95, 10))
    torch.cuda.synchronize()
    assert correct.shape == (3, 300, 10) and polyn.shape == (3, 300, 8) and torch.isfinite(hbbn[0, :rows[0]]).all()
