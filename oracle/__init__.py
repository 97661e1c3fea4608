"""CPU oracle for the yolov5_obb hot path — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (yolov5_obb_b200) never does, and has no CPU fallback.

Contents
  liboracle.so (obb_oracle.cpp)  rotated IoU + greedy rotated NMS, scalar C++, no FMA
  postprocess.py                 non_max_suppression_obb restated with torch CPU ops + the C++ NMS
  detect_ref.py / loss_ref.py / model_ref.py   fp32 torch restatements of Detect decode, ComputeLoss,
                                 and the Conv/C3/SPPF graph (floating-point kernels keep a torch fp32
                                 reference, as the task allows)
  build_ref.py                   compiles the reference's own kernels into oracle/_ref (pinning)
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build() -> None:
    """Compile liboracle.so with g++ (seconds)."""
    subprocess.check_call(["make", "-s", "-C", str(_HERE), "liboracle.so"])


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "obb_oracle.cpp").stat().st_mtime:
            build()
        L = ctypes.CDLL(str(so))
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.oracle_iou_pairs.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int]
        L.oracle_iou_pairs.restype = None
        L.oracle_nms_rotated.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_float, ctypes.c_int, ip]
        L.oracle_nms_rotated.restype = ctypes.c_int64
        L.oracle_obb_nms.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_float, ctypes.c_int, ip]
        L.oracle_obb_nms.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def iou_pairs(a, b, variant: int = 0) -> np.ndarray:
    """IoU of pairs (a[i], b[i]); boxes are (cx, cy, w, h, theta_rad).  variant 0 host hull, 1 device hull."""
    a, b = _f32(a).reshape(-1, 5), _f32(b).reshape(-1, 5)
    out = np.empty(a.shape[0], np.float32)
    lib().oracle_iou_pairs(_fp(a), _fp(b), _fp(out), a.shape[0], variant)
    return out


def nms_rotated(dets, scores, thr: float, mode: int = 1) -> np.ndarray:
    """Keep indices (score-descending).  mode 0 = reference CPU (>=), mode 1 = reference CUDA (>)."""
    d, s = _f32(dets).reshape(-1, 5), _f32(scores).reshape(-1)
    keep = np.empty(d.shape[0], np.int64)
    n = lib().oracle_nms_rotated(_fp(d), _fp(s), d.shape[0], thr, mode,
                                 keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return keep[:n].copy()


def obb_nms(dets, scores, thr: float, mode: int = 1) -> np.ndarray:
    """nms_rotated_wrapper.obb_nms index semantics (drops min(w,h) < 0.001 first)."""
    d, s = _f32(dets).reshape(-1, 5), _f32(scores).reshape(-1)
    keep = np.empty(max(d.shape[0], 1), np.int64)
    n = lib().oracle_obb_nms(_fp(d), _fp(s), d.shape[0], thr, mode,
                             keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return keep[:n].copy()
