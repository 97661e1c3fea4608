"""GPU parity of the fused post-NMS geometry + validation matching kernel (csrc/val_match.cu) against outputs
of the REFERENCE functions chained as val.py:226-250 does (tests/golden/valmatch_golden.npz): `correct` matrices equal,
native-space polygons and HBB boxes within 2 ulp of the ATen chain (every step is a separately rounded fp32 op in the same
order; only cos / sin implementations differ in the last bit)."""
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
        # the golden ran on the CPU; torch's CPU and CUDA cos / sin and this kernel's cosf / sinf differ from one another in the last
        # ulp, so geometry is held to 1e-3 px against the golden and to 2 ulp (2e-4 px at 1000 px) against the same ATen chain
        # evaluated on this GPU below; the threshold decisions (`correct`) must be equal
        np.testing.assert_allclose(polyn[si, :n].cpu().numpy(), G[f"{seed}/{si}/polyn"], rtol=0, atol=1e-3)
        np.testing.assert_allclose(hbbn[si, :n].cpu().numpy(), G[f"{seed}/{si}/hbbn"], rtol=0, atol=1e-3)
        pred = torch.from_numpy(dets[si, :n]).to(DEV)
        gain, (px, py) = shapes[si][1][0][0], shapes[si][1][1]
        c, w, h, th = pred[:, :2], pred[:, 2:3], pred[:, 3:4], pred[:, 4:5]          # utils/rboxs_utils.py:106-126 (torch branch)
        Cos, Sin = torch.cos(th), torch.sin(th)
        v1, v2 = torch.cat((w / 2 * Cos, -w / 2 * Sin), -1), torch.cat((-h / 2 * Sin, -h / 2 * Cos), -1)
        poly = torch.cat((c + v1 + v2, c + v1 - v2, c - v1 - v2, c - v1 + v2), -1)
        poly[:, 0::2] -= px                                                          # scale_polys (utils/general.py:636-650)
        poly[:, 1::2] -= py
        poly /= gain
        assert torch.allclose(polyn[si, :n], poly, rtol=0, atol=2e-4), (si, (polyn[si, :n] - poly).abs().max().item())
        x, y = poly[:, 0::2], poly[:, 1::2]                                          # poly2hbb + xywh2xyxy
        xc, yc = (x.amax(1) + x.amin(1)) / 2.0, (y.amax(1) + y.amin(1)) / 2.0
        ww, hh = x.amax(1) - x.amin(1), y.amax(1) - y.amin(1)
        box = torch.stack((xc - ww / 2, yc - hh / 2, xc + ww / 2, yc + hh / 2), 1)
        assert torch.allclose(hbbn[si, :n], box, rtol=0, atol=2e-4), (si, (hbbn[si, :n] - box).abs().max().item())


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
