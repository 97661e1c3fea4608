"""ctypes handles for the training-only kernels (BatchNorm backward, layout copies, tensor-core wgrad)."""
import ctypes
from ctypes import c_void_p, c_int, c_int64

import torch

from . import _lib


class WgradDesc(ctypes.Structure):
    """Mirror of y5obb_wgrad_desc (include/y5obb.h)."""
    _fields_ = [("dz", c_void_p), ("dz_pix_stride", c_int64), ("x", c_void_p), ("x_pix_stride", c_int64),
                ("x_row_stride", c_int64), ("x_img_stride", c_int64), ("dw", c_void_p),
                ("dw_tap_stride", c_int64), ("dw_co_stride", c_int64), ("dw_ci_stride", c_int64),
                ("B", c_int), ("Cout", c_int), ("Ho", c_int), ("Wo", c_int), ("Cin", c_int), ("Hi", c_int), ("Wi", c_int),
                ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad_h", c_int), ("pad_w", c_int),
                ("co_group", c_int), ("co_group_pad", c_int)]


class Wgrad:
    """dW += sum_pixels dz * x over fixed NHWC buffers (TMA descriptors baked at creation).
    dz / x: (ptr, pix_stride) of the channel slice; `keep` holds the owning tensors alive.
    param_layout=False: dw is [KH*KW][Cout][Cin]; True: dw is the nn.Conv2d parameter layout [Cout][Cin][KH][KW].
    co_group=(real, padded): Detect's per-anchor padded channel groups (rows of padding channels are skipped).
    x_strides=(row, image) in elements: explicit strides of x (allows overlapping pixel windows, x_pix_stride < Cin)."""

    def __init__(self, dz_ptr: int, dz_pix_stride: int, x_ptr: int, x_pix_stride: int, dw: torch.Tensor, B, Cout, Ho, Wo,
                 Cin, Hi, Wi, k, stride, pad, keep=(), param_layout=False, co_group=(0, 0), x_strides=(0, 0)):
        kh, kw = (k, k) if isinstance(k, int) else k
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        st = (1, Cin * kh * kw, kh * kw) if param_layout else (0, 0, 0)
        d = WgradDesc(dz_ptr, dz_pix_stride, x_ptr, x_pix_stride, x_strides[0], x_strides[1], dw.data_ptr(), st[0], st[1], st[2],
                      B, Cout, Ho, Wo, Cin, Hi,
                      Wi, kh, kw, stride, ph, pw, co_group[0], co_group[1])
        rows = Cout if not co_group[1] else Cout // co_group[1] * co_group[0]
        assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == kh * kw * rows * Cin
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


PACK_FWD, PACK_DGRAD, PACK_STEM, PACK_DETECT, PACK_DETECT_DGRAD, PACK_DETECT_BIAS, PACK_DGRAD_S2 = range(7)


class PackEntry(ctypes.Structure):
    """Mirror of y5obb_pack_entry (include/y5obb.h)."""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("kind", c_int), ("Cout", c_int), ("Cin", c_int), ("KH", c_int),
                ("KW", c_int), ("rows_pad", c_int), ("cols_pad", c_int), ("group_real", c_int), ("group_pad", c_int)]


class PackPlan:
    """All layers' fp32 parameters -> packed bf16 operand buffers in one launch (after every optimiser step).
    entries: (kind, src parameter tensor fp32, dst packed tensor, group_real, group_pad); shapes are read off the tensors."""

    def __init__(self, entries, device):
        arr = (PackEntry * len(entries))()
        self._keep = []
        for i, (kind, src, dst, g_real, g_pad) in enumerate(entries):
            assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
            if kind == PACK_DETECT_BIAS:
                cout, cin, kh, kw = src.shape[0], 1, 1, 1
                rows_pad, cols_pad = 1, dst.numel()
                assert dst.dtype == torch.float32
            else:
                cout, cin, kh, kw = src.shape
                assert dst.dtype == torch.bfloat16 and dst.dim() == 3
                rows_pad, cols_pad = dst.shape[1], dst.shape[2]
                taps = 3 if kind == PACK_STEM else ((2 if g_real >> 1 else 1) * (2 if g_real & 1 else 1) if kind == PACK_DGRAD_S2 else kh * kw)
                assert dst.shape[0] == taps
            arr[i] = PackEntry(src.data_ptr(), dst.data_ptr(), kind, cout, cin, kh, kw, rows_pad, cols_pad, g_real, g_pad)
            self._keep.append((src, dst))
        self._h = c_void_p()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().y5obb_pack_plan_create(arr, len(entries), ctypes.byref(self._h))
        _lib.check(rc, "y5obb_pack_plan_create")

    def run(self, stream=None):
        with torch.cuda.device(self.device):
            rc = _lib.lib().y5obb_pack_plan_run(self._h, stream if stream is not None else _lib.stream_ptr(self.device))
        _lib.check(rc, "y5obb_pack_plan_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().y5obb_pack_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class SgdEntry(ctypes.Structure):
    """Mirror of y5obb_sgd_entry (include/y5obb.h)."""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("mom", c_void_p), ("ema", c_void_p), ("n", c_int64),
                ("weight_decay", ctypes.c_float), ("group", c_int)]


class FusedSGDEMA:
    """optimizer.step() + ema.update() of train.py:336-342 as ONE launch (tests/test_sgd_ema_gpu.py: equal to torch.optim.SGD
    with the reference's parameter groups + ModelEMA within 1e-5 relative over three steps).

    groups: (g0, g1, g2) parameter lists as train.py:148-156 builds them; grads: {parameter: fp32 gradient tensor} with FIXED
    storage (e.g. views of the backward plan's flat buffer); ema_model: a deep copy whose state_dict mirrors model's.
    step(lr=(lr0, lr1, lr2), momentum, ema_decay) applies torch.optim.SGD(nesterov=True) arithmetic and the EMA rule."""

    def __init__(self, model, groups, grads, ema_model=None, weight_decay: float = 0.0, momentum_buffers=None):
        self.device = next(model.parameters()).device
        ema_sd = dict(ema_model.state_dict()) if ema_model is not None else {}
        name_of = {id(p): n for n, p in model.named_parameters()}
        self.mom, ent, self._keep = {}, [], []
        for gi, plist in enumerate(groups):
            for p in plist:
                g = grads[p]
                assert p.dtype == torch.float32 and g.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
                m = momentum_buffers[p] if momentum_buffers is not None else torch.zeros_like(p)
                self.mom[p] = m
                e = ema_sd.get(name_of[id(p)])
                ent.append(SgdEntry(p.data_ptr(), g.data_ptr(), m.data_ptr(), e.data_ptr() if e is not None else None, p.numel(),
                                    float(weight_decay) if gi == 1 else 0.0, gi))
                self._keep += [p, g, m, e]
        if ema_model is not None:  # floating-point buffers (BatchNorm running statistics): EMA only
            msd = dict(model.state_dict())
            pnames = set(name_of.values())
            for k, e in ema_sd.items():
                if k in pnames or not e.dtype.is_floating_point:
                    continue
                v = msd[k]
                ent.append(SgdEntry(v.data_ptr(), None, None, e.data_ptr(), v.numel(), 0.0, -1))
                self._keep += [v, e]
        arr = (SgdEntry * len(ent))(*ent)
        self._h = c_void_p()
        with torch.cuda.device(self.device):
            rc = _lib.lib().y5obb_sgd_ema_plan_create(arr, len(ent), ctypes.byref(self._h))
        _lib.check(rc, "y5obb_sgd_ema_plan_create")
        self.steps = 0 if momentum_buffers is None else 1  # adopted buffers already hold a first step

    def step(self, lr, momentum: float, ema_decay: float = 0.0, stream=None):
        lr3 = (ctypes.c_float * 3)(*[float(v) for v in lr])
        with torch.cuda.device(self.device):
            rc = _lib.lib().y5obb_sgd_ema_plan_run(self._h, lr3, float(momentum), float(ema_decay), int(self.steps == 0),
                                                   stream if stream is not None else _lib.stream_ptr(self.device))
        _lib.check(rc, "y5obb_sgd_ema_plan_run")
        self.steps += 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().y5obb_sgd_ema_plan_destroy(h)
            except Exception:
                pass
            self._h = None
