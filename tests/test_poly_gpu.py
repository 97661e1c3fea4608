"""GPU parity of the tile-merge polygon NMS (csrc/poly_nms.cu) against the reference's polyiou.cpp outputs
(tests/golden/poly_golden.npz): IoU bit-equal, keep lists equal."""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
DEV = "cuda:0"


def test_iou_poly_bit_equal():
    from yolov5_obb_b200.devkit import iou_poly_pairs
    G = np.load(ROOT / "tests" / "golden" / "poly_golden.npz")
    got = iou_poly_pairs(torch.from_numpy(G["P"]).to(DEV), torch.from_numpy(G["Q"]).to(DEV)).cpu().numpy()
    assert np.array_equal(got, G["iou"]), np.abs(got - G["iou"]).max()


def test_tile_merge_nms_keep_lists():
    from yolov5_obb_b200.devkit import py_cpu_nms_poly_fast
    G = np.load(ROOT / "tests" / "golden" / "poly_golden.npz")
    for k in range(3):
        keep = py_cpu_nms_poly_fast(torch.from_numpy(G[f"dets{k}"]).to(DEV), float(G[f"thr{k}"]))
        assert keep.cpu().tolist() == G[f"keep{k}"].tolist()
