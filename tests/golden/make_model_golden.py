"""Golden vectors from the REFERENCE model (models/yolo.py Model, eval forward) — run in the authoring
container:  python tests/golden/make_model_golden.py
Writes model_golden.npz: for yolov5n and yolov5s (nc=15, torch seed 0, tests/modelgen.seeded_state), the
eval output [1, A, 200] on a seeded 1x3x64x96 image, plus a few intermediate checksums."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_import  # noqa: E402

ref_import.setup()
sys.path.insert(0, str(HERE.parents[1]))
from models.yolo import Model  # noqa: E402  (the reference)
from tests.modelgen import seeded_state  # noqa: E402


def main():
    out = {}
    for size in ("n", "s"):
        torch.manual_seed(0)
        m = Model(f"models/yolov5{size}.yaml", ch=3, nc=15)
        seeded_state(m, 0).eval()
        g = torch.Generator().manual_seed(123)
        x = torch.rand(1, 3, 64, 96, generator=g)
        with torch.no_grad():
            pred, raw = m(x)
        out[f"{size}/x"] = x.numpy()
        out[f"{size}/pred"] = pred.numpy()
        out[f"{size}/param_sum"] = np.float64(sum(p.double().sum().item() for p in m.parameters()))
        # fused path must agree with itself (fuse_conv_and_bn)
        m.fuse()
        with torch.no_grad():
            pred_f, _ = m(x)
        out[f"{size}/pred_fused_maxdiff"] = np.float32((pred_f - pred).abs().max().item())
        print(size, pred.shape, float(pred.abs().mean()), out[f"{size}/pred_fused_maxdiff"])
    np.savez_compressed(HERE / "model_golden.npz", **out)


if __name__ == "__main__":
    main()
