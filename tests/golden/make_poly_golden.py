"""Golden vectors for the devkit's polygon IoU: the reference's DOTA_devkit/polyiou.cpp compiled in place
(oracle/build_ref.py build_polyiou) on seeded quadrilateral pairs, plus keep lists of the tile-merge NMS
(ResultMerge_multi_process.py:62-123 restated in oracle/poly_ref.py, driven by the REFERENCE IoU).
python tests/golden/make_poly_golden.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle.build_ref import build_polyiou, load_polyiou  # noqa: E402
from oracle import poly_ref  # noqa: E402
from tests.polygen import quad_pairs, merge_dets  # noqa: E402


def main():
    assert build_polyiou()
    lib = load_polyiou()
    P, Q = quad_pairs(seed=0)
    iou = np.zeros(len(P))
    lib.ref_iou_poly_pairs(P.ctypes.data, Q.ctypes.data, iou.ctypes.data, len(P))

    def ref_iou(a, b):
        o = np.zeros(1)
        a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
        lib.ref_iou_poly_pairs(a.ctypes.data, b.ctypes.data, o.ctypes.data, 1)
        return o[0]

    out = {"P": P, "Q": Q, "iou": iou}
    for k, (n, thr) in enumerate([(300, 0.2), (800, 0.1), (500, 0.5)]):
        D = merge_dets(n, seed=10 + k)
        out[f"dets{k}"] = D
        out[f"thr{k}"] = np.float64(thr)
        out[f"keep{k}"] = np.array(poly_ref.py_cpu_nms_poly_fast(D, thr, iou=ref_iou), np.int64)
    np.savez_compressed(HERE / "poly_golden.npz", **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
