"""CPU test: rbox utility restatements against outputs of the REFERENCE functions (rbox_golden.npz)."""
from pathlib import Path

import numpy as np

from oracle import rbox_ref

G = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / "rbox_golden.npz")


def test_rbox2poly_poly2hbb_scale_polys():
    polys = rbox_ref.rbox2poly(G["rboxes"])
    np.testing.assert_allclose(polys, G["polys"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(rbox_ref.poly2hbb(G["polys"]), G["hbb"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(rbox_ref.scale_polys((1024, 1024), G["polys"], (1689, 2425)), G["scaled"], rtol=1e-6, atol=1e-4)


def test_gaussian_label_quirks():
    for key, sig in (("csl2", 2.0), ("csl6", 6.0)):
        rows = np.stack([rbox_ref.gaussian_label_cpu(a, 180, 0, sig) for a in G["angles"]]).astype(np.float32)
        assert np.array_equal(rows, G[key])
    # peak = 90 - trunc(90 - angle): 0.4->1, 45.7->46, 90.3->90, 135.5->135, 179.5->179
    peaks = G["csl2"].argmax(1)[:5].tolist()
    assert peaks == [1, 46, 90, 135, 179]
    assert abs(G["csl2"][0].sum() - 5.013257) < 1e-5
