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


# ---- the devkit's GPU entry points (DOTA_devkit/poly_nms_gpu), float arithmetic ------------------------------------------
def poly_gpu_nms(dets, thresh: float, device_id: int = 0):
    """DOTA_devkit/poly_nms_gpu/poly_nms.pyx:10-27: dets float32 numpy [n, 9] (8 coordinates + score) -> list of kept indices.
    Sorts by descending score on the host exactly as the .pyx does (`scores.argsort()[::-1]`), then calls the library's
    `_poly_nms` entry point (host pointers, the reference's own C ABI)."""
    import ctypes
    import numpy as np
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n, dim = dets.shape
    order = dets[:, 8].argsort()[::-1]
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(max(n, 1), dtype=np.int32)
    num = ctypes.c_int(0)
    _lib.lib().y5obb_devkit_poly_nms(keep.ctypes.data, ctypes.addressof(num), sorted_dets.ctypes.data, n, dim, float(thresh),
                                     int(device_id))
    return list(order[keep[:num.value]])


def poly_overlaps(boxes, query_boxes, device_id: int = 0):
    """DOTA_devkit/poly_nms_gpu/poly_overlaps.pyx: float32 numpy [n, 5] x [k, 5] rotated boxes (cx, cy, w, h, angle) ->
    [n, k] IoU matrix, through the library's `_overlaps` entry point."""
    import numpy as np
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    q = np.ascontiguousarray(query_boxes, dtype=np.float32)
    out = np.zeros((b.shape[0], q.shape[0]), dtype=np.float32)
    if out.size:
        _lib.lib().y5obb_devkit_overlaps(out.ctypes.data, b.ctypes.data, q.ctypes.data, b.shape[0], q.shape[0], int(device_id))
    return out


def poly_overlaps_device(boxes: torch.Tensor, query_boxes: torch.Tensor) -> torch.Tensor:
    """The same IoU matrix for CUDA tensors, asynchronous on the current stream (no host round trip)."""
    _lib.require_cuda(boxes, "boxes")
    _lib.require_cuda(query_boxes, "query_boxes")
    b, q = boxes.reshape(-1, 5).float().contiguous(), query_boxes.reshape(-1, 5).float().contiguous()
    out = torch.empty((b.shape[0], q.shape[0]), dtype=torch.float32, device=b.device)
    with torch.cuda.device(b.device):
        rc = _lib.lib().y5obb_poly_overlaps_f32(b.data_ptr(), q.data_ptr(), b.shape[0], q.shape[0], out.data_ptr(),
                                                _lib.stream_ptr(b.device))
    _lib.check(rc, "y5obb_poly_overlaps_f32")
    return out
