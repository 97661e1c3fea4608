"""Golden vectors for the validation matching block, produced by the REFERENCE functions themselves (val.py:69-92 process_batch,
utils/rboxs_utils.py rbox2poly / poly2hbb, utils/general.py xywh2xyxy / scale_polys / scale_coords) run exactly as val.py:226-250
chains them, on the seeded batch of tests/valgen.py.  Only outputs are stored.  python tests/golden/make_valmatch_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_import  # noqa: E402

ref_import.setup()
sys.path.insert(0, str(HERE.parents[1]))
import val as ref_val  # noqa: E402
from utils.general import xywh2xyxy, scale_polys, scale_coords  # noqa: E402
from utils.rboxs_utils import rbox2poly, poly2hbb  # noqa: E402
from tests.valgen import synth_val_batch  # noqa: E402


def main():
    out = {}
    for seed in (0, 1):
        dets, counts, targets, shapes = synth_val_batch(seed)
        iouv = torch.linspace(0.5, 0.95, 10)
        targets = torch.from_numpy(targets)
        for si in range(dets.shape[0]):
            pred = torch.from_numpy(dets[si, :counts[si]]).clone()
            labels = targets[targets[:, 0] == si, 1:7]                       # val.py:214
            shape = shapes[si][0]
            im_shape = (1024, 1024)
            poly = rbox2poly(pred[:, :5])                                     # val.py:227
            pred_poly = torch.cat((poly, pred[:, -2:]), dim=1)
            pred_polyn = pred_poly.clone()
            scale_polys(im_shape, pred_polyn[:, :8], shape, shapes[si][1])    # val.py:233
            hbboxn = xywh2xyxy(poly2hbb(pred_polyn[:, :8]))
            pred_hbbn = torch.cat((hbboxn, pred_polyn[:, -2:]), dim=1)
            if len(labels):
                tpoly = rbox2poly(labels[:, 1:6])
                tbox = xywh2xyxy(poly2hbb(tpoly))
                scale_coords(im_shape, tbox, shape, shapes[si][1])
                labels_hbbn = torch.cat((labels[:, 0:1], tbox), 1)
                correct = ref_val.process_batch(pred_hbbn, labels_hbbn, iouv)
            else:
                correct = torch.zeros(pred.shape[0], 10, dtype=torch.bool)
            out[f"{seed}/{si}/correct"] = correct.numpy()
            out[f"{seed}/{si}/polyn"] = pred_polyn[:, :8].numpy()
            out[f"{seed}/{si}/hbbn"] = hbboxn.numpy()
            print(seed, si, tuple(correct.shape), int(correct[:, 0].sum()), int(correct[:, -1].sum()))
    np.savez_compressed(HERE / "valmatch_golden.npz", **out)


if __name__ == "__main__":
    main()
