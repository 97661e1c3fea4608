"""CPU: the devkit polygon IoU / tile-merge NMS restatement (oracle/poly_ref.py, groundwork for SURVEY section 8f rank 4)
against the reference's own DOTA_devkit/polyiou.cpp compiled in place (tests/golden/poly_golden.npz; and live against
oracle/_ref/libref_polyiou.so where it exists): IoU values bit-equal (same double arithmetic, no FMA), keep lists equal."""
from pathlib import Path

import numpy as np

from oracle import poly_ref

ROOT = Path(__file__).resolve().parents[1]
G = np.load(ROOT / "tests" / "golden" / "poly_golden.npz")


def test_iou_poly_bit_equal_to_reference():
    P, Q, want = G["P"], G["Q"], G["iou"]
    got = np.array([poly_ref.iou_poly(p, q) for p, q in zip(P, Q)])
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert (want > 0).mean() > 0.4 and np.nanmax(want) <= 1.0 + 1e-9
    # identical boxes (any winding / start corner) give 1 up to rounding
    same = np.arange(len(P)) % 6
    assert np.abs(want[(same == 1) | (same == 2)] - 1.0).max() < 1e-9
    so = ROOT / "oracle" / "_ref" / "libref_polyiou.so"
    if so.exists():   # the compiled reference itself, where present
        from oracle.build_ref import load_polyiou
        lib = load_polyiou()
        live = np.zeros(len(P))
        lib.ref_iou_poly_pairs(P.ctypes.data, Q.ctypes.data, live.ctypes.data, len(P))
        assert np.array_equal(live, want)


def test_tile_merge_nms_keep_lists():
    for k in range(3):
        D, thr = G[f"dets{k}"], float(G[f"thr{k}"])
        keep = poly_ref.py_cpu_nms_poly_fast(D, thr)
        assert keep == G[f"keep{k}"].tolist()
        sc = D[keep, 8]
        assert np.all(np.diff(sc) < 0) and len(set(keep)) == len(keep)     # descending score, no repeats
        for a in range(min(len(keep), 40)):                                 # survivors do not overlap above the threshold
            for b in range(a + 1, min(len(keep), 40)):
                assert poly_ref.iou_poly(D[keep[a], :8], D[keep[b], :8]) <= thr + 1e-12
