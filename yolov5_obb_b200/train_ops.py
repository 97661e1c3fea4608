"""ctypes handles for the training-only kernels (BatchNorm backward, layout copies, tensor-core wgrad)."""
import ctypes
from ctypes import c_void_p, c_int, c_int64

import torch

from . import _lib


class WgradDesc(ctypes.Structure):
    """Mirror of y5obb_wgrad_desc (include/y5obb.h)."""
    _fields_ = [("dz", c_void_p), ("dz_pix_stride", c_int64), ("x", c_void_p), ("x_pix_stride", c_int64), ("dw", c_void_p),
                ("B", c_int), ("Cout", c_int), ("Ho", c_int), ("Wo", c_int), ("Cin", c_int), ("Hi", c_int), ("Wi", c_int),
                ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad_h", c_int), ("pad_w", c_int)]


class Wgrad:
    """dW[tap][co][ci] += sum_pixels dz * x over fixed NHWC buffers (TMA descriptors baked at creation).
    dz / x: (ptr, pix_stride) of the channel slice; `keep` holds the owning tensors alive."""

    def __init__(self, dz_ptr: int, dz_pix_stride: int, x_ptr: int, x_pix_stride: int, dw: torch.Tensor, B, Cout, Ho, Wo,
                 Cin, Hi, Wi, k, stride, pad, keep=()):
        kh, kw = (k, k) if isinstance(k, int) else k
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        d = WgradDesc(dz_ptr, dz_pix_stride, x_ptr, x_pix_stride, dw.data_ptr(), B, Cout, Ho, Wo, Cin, Hi, Wi, kh, kw, stride,
                      ph, pw)
        assert dw.dtype == torch.float32 and dw.numel() == kh * kw * Cout * Cin
        self._keep = (dw,) + tuple(keep)
        self._h = c_void_p()
        self.device = dw.device
        with torch.cuda.device(self.device):
            rc = _lib.lib().y5obb_wgrad_create(ctypes.byref(d), ctypes.byref(self._h))
        _lib.check(rc, "y5obb_wgrad_create")

    def run(self, stream=None):
        rc = _lib.lib().y5obb_wgrad_run(self._h, stream if stream is not None else _lib.stream_ptr(self.device))
        _lib.check(rc, "y5obb_wgrad_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().y5obb_wgrad_destroy(h)
            except Exception:
                pass
            self._h = None


def nhwc_to_nchw(src_ptr: int, src_pix_stride: int, dst: torch.Tensor, B: int, C: int, H: int, W: int, phase_split=False,
                 stream=None):
    rc = _lib.lib().y5obb_nhwc_to_nchw(src_ptr, src_pix_stride, dst.data_ptr(), B, C, H * W, W if phase_split else 0,
                                       stream if stream is not None else _lib.stream_ptr(dst.device))
    _lib.check(rc, "y5obb_nhwc_to_nchw")
