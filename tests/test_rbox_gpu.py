"""GPU parity for the post-NMS geometry / CSL kernels against the REFERENCE's outputs (rbox_golden.npz)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / "rbox_golden.npz")
DEV = "cuda:0"


def test_rbox2poly_poly2hbb_scale_polys():
    from yolov5_obb_b200.rboxs_utils import rbox2poly, poly2hbb, scale_polys
    r = torch.from_numpy(G["rboxes"]).to(DEV)
    polys = rbox2poly(r)
    # cosf/sinf on the device vs the reference's CPU libm: <= 1 ulp of the trig value times the box size
    np.testing.assert_allclose(polys.cpu().numpy(), G["polys"], rtol=0, atol=2e-4)
    # same arithmetic as a chain of torch CUDA elementwise ops -> bit-exact against that chain
    th = r[:, 4:5]
    v1 = torch.cat((r[:, 2:3] / 2 * torch.cos(th), -r[:, 2:3] / 2 * torch.sin(th)), -1)
    v2 = torch.cat((-r[:, 3:4] / 2 * torch.sin(th), -r[:, 3:4] / 2 * torch.cos(th)), -1)
    c = r[:, :2]
    chain = torch.cat((c + v1 + v2, c + v1 - v2, c - v1 - v2, c - v1 + v2), -1)
    assert torch.equal(polys, chain)
    gp = torch.from_numpy(G["polys"]).to(DEV)
    assert np.array_equal(poly2hbb(gp).cpu().numpy(), G["hbb"])
    sp = scale_polys((1024, 1024), gp.clone(), (1689, 2425))
    np.testing.assert_allclose(sp.cpu().numpy(), G["scaled"], rtol=1e-6, atol=1e-4)
    assert rbox2poly(r.view(5, 100, 5)).shape == (5, 100, 8)
    with pytest.raises(RuntimeError):
        rbox2poly(r.cpu())


def test_gaussian_label_matches_reference_rows():
    from yolov5_obb_b200.rboxs_utils import gaussian_label
    a = torch.from_numpy(G["angles"]).to(DEV)
    for key, sig in (("csl2", 2.0), ("csl6", 6.0)):
        got = gaussian_label(a, 180, 0, sig).cpu().numpy()
        np.testing.assert_allclose(got, G[key], rtol=1e-6, atol=1e-12)
        assert np.array_equal(got.argmax(1), G[key].argmax(1))
