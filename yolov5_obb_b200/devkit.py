"""Host-side mirror of the DOTA devkit's tile-merge NMS (DOTA_devkit/ResultMerge_multi_process.py:62-123
py_cpu_nms_poly_fast over polyiou.iou_poly), backed by csrc/poly_nms.cu.

tests/test_poly_gpu.py: IoU bit-equal to the reference's polyiou.cpp, keep lists equal (fp64, no FMA contraction)."""
import torch

from . import _lib


def iou_poly_pairs(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """[n, 8] x [n, 8] polygons (x1 y1 ... x4 y4) on a CUDA device -> [n] fp64 IoU, as polyiou.iou_poly pair by pair."""
    _lib.require_cuda(p, "p")
    _lib.require_cuda(q, "q")
    p8, q8 = p.reshape(-1, 8).double().contiguous(), q.reshape(-1, 8).double().contiguous()
    assert p8.shape == q8.shape
    out = torch.empty(p8.shape[0], dtype=torch.float64, device=p8.device)
    with torch.cuda.device(p8.device):
        rc = _lib.lib().y5obb_poly_iou_pairs_f64(p8.data_ptr(), q8.data_ptr(), out.data_ptr(), p8.shape[0],
                                                 _lib.stream_ptr(p8.device))
    _lib.check(rc, "y5obb_poly_iou_pairs_f64")
    return out


def py_cpu_nms_poly_fast(dets: torch.Tensor, thresh: float) -> torch.Tensor:
    """dets [n, 9] (8 polygon coordinates + score) on a CUDA device -> int64 keep indices in descending score order
    (the reference returns the same list as Python ints)."""
    _lib.require_cuda(dets, "dets")
    d = dets.reshape(-1, 9).double().contiguous()
    n = d.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=d.device)
    nk = torch.zeros(1, dtype=torch.int64, device=d.device)
    L = _lib.lib()
    with torch.cuda.device(d.device):
        ws = _lib.workspace(L.y5obb_poly_nms_workspace_bytes(n), d.device, "poly_nms")
        rc = L.y5obb_poly_nms_f64(d.data_ptr(), n, float(thresh), keep.data_ptr(), nk.data_ptr(), ws.data_ptr(), ws.numel(),
                                  _lib.stream_ptr(d.device))
    _lib.check(rc, "y5obb_poly_nms_f64")
    return keep[:int(nk.item())]
