"""Golden vectors from the REFERENCE's utils/rboxs_utils.py (torch branches) and general.scale_polys.
python tests/golden/make_rbox_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_import  # noqa: E402

ref_import.setup()
sys.path.insert(0, str(HERE.parents[1]))
from utils.rboxs_utils import rbox2poly, poly2hbb, gaussian_label_cpu, poly2rbox  # noqa: E402
from utils.general import scale_polys  # noqa: E402
from tests.boxgen import rboxes  # noqa: E402


def main():
    d, _, _ = rboxes(500, 1024, 3, class_offset=False, theta_grid=False)
    t = torch.from_numpy(d)
    polys = rbox2poly(t)
    out = {"rboxes": d, "polys": polys.numpy(), "hbb": poly2hbb(polys).numpy()}
    out["scaled"] = scale_polys((1024, 1024), polys.clone(), (1689, 2425)).numpy()
    angles = np.array([0.4, 45.7, 90.3, 135.5, 179.5, 0.0, 89.999, 90.0, 45.00000000000001, 179.99, 12.0, 91.2])
    out["angles"] = angles
    out["csl2"] = np.stack([gaussian_label_cpu(a, 180, u=0, sig=2.0) for a in angles]).astype(np.float32)
    out["csl6"] = np.stack([gaussian_label_cpu(a, 180, u=0, sig=6.0) for a in angles]).astype(np.float32)
    # poly2rbox known answers (cv2.minAreaRect 4.13)
    P = np.array([[0, 0, 100, 0, 100, 50, 0, 50], [0, 0, 50, 0, 50, 100, 0, 100],
                  [1707.0, 1539.0, 1683.0, 1523.0, 1689.0, 1513.0, 1713.0, 1529.0]], np.float64)
    out["p2r_polys"] = P
    out["p2r_rboxes"] = poly2rbox(P, use_pi=True)
    np.savez_compressed(HERE / "rbox_golden.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
