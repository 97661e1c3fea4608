"""GPU parity for the post-NMS geometry / CSL kernels against the REFERENCE's outputs (rbox_golden.npz)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / "rbox_golden.npz")
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parents[1]


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


def test_poly2rbox_matches_oracle_and_reference_outputs():
    """k_poly2rbox (A8): identical to the float64 oracle restatement (same hull, same edge choice, <= 1e-9) and, through
    it, to the REFERENCE poly2rbox outputs (cv2.minAreaRect 4.13.0) up to the ties cv2's float32 rounding decides
    (tests/test_oracle_p2r.py states the tolerances)."""
    import numpy as np
    from oracle import rbox_ref
    from tests.test_oracle_p2r import compare_p2r
    from yolov5_obb_b200.rboxs_utils import poly2rbox
    G = np.load(ROOT / "tests" / "golden" / "p2r_golden.npz")
    P = G["polys"]
    t = torch.from_numpy(P).to(DEV)
    got = poly2rbox(t, use_pi=True)
    assert got.dtype == torch.float64 and got.shape == (len(P), 5)
    want = rbox_ref.poly2rbox(P, use_pi=True)
    g = got.cpu().numpy()
    assert np.abs(g[:, :4] - want[:, :4]).max() < 1e-9
    d = np.abs(g[:, 4] - want[:, 4])
    assert np.minimum(d, rbox_ref.PI_REF - d).max() < 1e-9        # (the wrap point -pi/2 == +pi/2)
    ties = compare_p2r(g, G["rbox_pi"], P)
    assert ties < 0.15 * len(P)
    deg = poly2rbox(t, use_pi=False).cpu().numpy()
    assert np.abs(deg[:, 4] - (g[:, 4] * 180 / rbox_ref.PI_REF + 90)).max() < 1e-9 and deg[:, 4].min() >= 0 and deg[:, 4].max() < 180 + 1e-9
    # with CSL rows, as the dataloader asks (utils/datasets.py:639-641): rows of the REFERENCE for non-tie polygons
    rb, csl = poly2rbox(t[:64], num_cls_thata=180, radius=2.0, use_pi=True, use_gaussian=True)
    ref_rb, ref_csl = G["csl_rbox"], G["csl"]
    same = np.abs(rb.cpu().numpy()[:, 4] - ref_rb[:, 4]) < 1e-4
    assert same.mean() > 0.8 and csl.shape == (64, 180)
    # the CSL peak bin follows the (float) angle: identical rows wherever the angle agrees to well within a bin
    ang = (ref_rb[:, 4] * 180 / rbox_ref.PI_REF + 90)
    safe = same & (np.abs(ang - np.round(ang)) > 1e-2)
    assert np.abs(csl.cpu().numpy()[safe] - ref_csl[safe]).max() < 1e-6
