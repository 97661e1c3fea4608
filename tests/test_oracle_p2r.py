"""CPU: the poly2rbox restatement (oracle/rbox_ref.py, min-area rectangle by hull edges) against outputs of the
REFERENCE function, whose arithmetic is cv2.minAreaRect 4.13.0 (tests/golden/p2r_golden.npz, make_p2r_golden.py).
Tolerances (cv2 works in float32): centre / sides 2e-2 px + 2e-4 relative; theta 2e-3 rad, compared modulo pi/2 when
the rectangle is square within 1e-3 (the long-edge choice is then decided by cv2's float32 noise), modulo pi otherwise
(theta = -pi/2 and +pi/2 are the same orientation at the wrap)."""
from pathlib import Path

import numpy as np

from oracle import rbox_ref

ROOT = Path(__file__).resolve().parents[1]


def _encloses(rb, pts, slack):
    """all points inside the rectangle (cx, cy, l, s, theta) grown by `slack` px (long edge along (cos t, -sin t))"""
    c, s = np.cos(rb[4]), np.sin(rb[4])
    d = pts - rb[:2]
    a, b = d @ np.array([c, -s]), d @ np.array([s, c])
    return np.all(np.abs(a) <= rb[2] / 2 + slack) and np.all(np.abs(b) <= rb[3] / 2 + slack)


def compare_p2r(got, ref, polys):
    """Row-wise agreement; returns the number of TIES: polygons (typically integer-rounded DOTA boxes, which are
    parallelograms) for which two hull edges give enclosing rectangles of equal area in exact arithmetic, so that cv2's
    float32 rounding decides which one the reference returns.  A tie is accepted when both rectangles enclose the points
    and their areas agree to 1e-4 relative."""
    ties = 0
    for i, (g, r) in enumerate(zip(got, ref)):
        try:
            _compare_row(i, g, r, polys)
        except AssertionError:
            pts = np.asarray(polys[i], np.float64).reshape(4, 2)
            ag, ar = g[2] * g[3], r[2] * r[3]
            assert abs(ag - ar) <= 1e-4 * max(ar, 1.0) and _encloses(g, pts, 0.05) and _encloses(r, pts, 0.05), (i, g, r, polys[i])
            ties += 1
    return ties


def _compare_row(i, g, r, polys):
    if True:
        tol = 2e-2 + 2e-4 * max(abs(r[2]), 1.0)
        l, s = r[2], r[3]
        if s < 1e-6:            # degenerate input (collinear / repeated points): only centre and length are defined
            assert abs(g[0] - r[0]) < tol and abs(g[1] - r[1]) < tol and abs(g[2] - l) < tol and g[3] < tol, (i, g, r)
            return
        square = abs(l - s) <= 1e-3 * l + 2e-2
        assert abs(g[0] - r[0]) < tol and abs(g[1] - r[1]) < tol, (i, g, r, polys[i])
        if square:
            assert abs(g[2] - l) < 2 * tol + 1e-3 * l and abs(g[3] - s) < 2 * tol + 1e-3 * l, (i, g, r)
            period = rbox_ref.PI_REF / 2
        else:
            assert abs(g[2] - l) < tol and abs(g[3] - s) < tol, (i, g, r, polys[i])
            period = rbox_ref.PI_REF
        d = (g[4] - r[4]) % period
        d = min(d, period - d)
        # a thin rectangle's orientation is ill-conditioned in integer-rounded corners only through cv2's float32: scale by 1/l
        assert d < 2e-3 + 0.05 / max(l, 1.0), (i, g, r, polys[i])


def test_poly2rbox_oracle_matches_reference_outputs():
    G = np.load(ROOT / "tests" / "golden" / "p2r_golden.npz")
    P, R = G["polys"], G["rbox_pi"]
    got = rbox_ref.poly2rbox(P, use_pi=True)
    assert got.shape == R.shape == (len(P), 5)
    ties = compare_p2r(got, R, P)
    print("ties decided by cv2's float32 rounding:", ties, "of", len(P))
    assert ties < 0.15 * len(P)
    assert np.all(got[:, 4] >= -rbox_ref.PI_REF / 2 - 1e-12) and np.all(got[:, 4] < rbox_ref.PI_REF / 2)
    # the earlier known answers are part of the set
    K = np.load(ROOT / "tests" / "golden" / "rbox_golden.npz")
    compare_p2r(rbox_ref.poly2rbox(K["p2r_polys"]), K["p2r_rboxes"], K["p2r_polys"])
