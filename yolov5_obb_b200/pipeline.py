"""End-to-end detection pipeline from HOST images: the loop of /root/reference/detect.py:104-122 and
val.py:183-207 (pre-process, inference, NMS), software-pipelined two deep: the host->device copy of batch i+1 runs on
a second stream while batch i computes (two device input buffers, event-ordered), and the device->host copy of batch
i's detections is asynchronous, so the host turns batch i-1 into per-image tensors while the GPU works on batch i
(the stream is never drained inside the loop)."""
from typing import Iterable, Iterator, List

import torch

from .general import non_max_suppression_obb


class DetectPipeline:
    def __init__(self, model, conf_thres: float = 0.25, iou_thres: float = 0.45, max_det: int = 1500,
                 multi_label: bool = True, classes=None, agnostic: bool = False, device=None, fused_detect: bool = True):
        self.model = model
        # fused_detect: Model.detect_records + the post-process on the compact records (the [B, A, no] prediction tensor never
        # crosses HBM); False: Model.forward + non_max_suppression_obb on the tensor, as the two separate reference calls
        self.fused = bool(fused_detect) and hasattr(model, "detect_records")
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic,
                       multi_label=multi_label, max_det=max_det)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("DetectPipeline needs a CUDA device; there is no CPU path")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs = [None, None]
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _host_slot(self, slot, B, max_det):
        hs = getattr(self, "_hslots", None)
        if hs is None or hs[0][0].shape[0] != B or hs[0][0].shape[1] != max_det:
            self._hslots = hs = [(torch.empty((B, max_det, 7), dtype=torch.float32).pin_memory(),
                                  torch.empty(B + 1, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(2)]
        return hs[slot]

    def _upload(self, slot: int, x_host: torch.Tensor, first_use: bool) -> None:
        if self._bufs[slot] is None or self._bufs[slot].shape != x_host.shape or self._bufs[slot].dtype != x_host.dtype:
            self._bufs[slot] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
            first_use = True
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self._consumed[slot])  # the engine finished reading this buffer
            self._bufs[slot].copy_(x_host, non_blocking=True)
            self._copied[slot].record(self.copy_stream)
        self.h2d_bytes += x_host.numel() * x_host.element_size()

    def __call__(self, host_batches: Iterable[torch.Tensor]) -> Iterator[List[torch.Tensor]]:
        """host_batches: pinned uint8 (0..255) or float (0..1) tensors [B,3,H,W].  Yields, per batch, the list of
        per-image detections [n,7] (cx, cy, l, s, theta, conf, cls) as HOST tensors."""
        it = iter(host_batches)
        cur = next(it, None)
        if cur is None:
            return
        used = [False, False]
        self._upload(0, cur, True)
        used[0] = True
        i = 0
        pending = None
        compute = torch.cuda.current_stream(self.device)
        while cur is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                self._upload(slot ^ 1, nxt, not used[slot ^ 1])
                used[slot ^ 1] = True
            compute.wait_event(self._copied[slot])
            pred = self.model.detect_records(self._bufs[slot]) if self.fused else self.model(self._bufs[slot])[0]
            self._consumed[slot].record(compute)  # the first kernel has consumed the input by now (stream order)
            packed, counts, cap = non_max_suppression_obb(pred, return_packed="async", **self.kw)
            hout, hcnt, ev = self._host_slot(slot, packed.shape[0], packed.shape[1])
            hout.copy_(packed, non_blocking=True)   # D2H of this batch's result, one copy; nobody waits for it here
            hcnt.copy_(counts, non_blocking=True)
            ev.record(compute)
            self.d2h_bytes += packed.numel() * 4 + counts.numel() * 8
            if pending is not None:
                yield self._finish(pending)
            pending = (hout, hcnt, ev, cap, cur)
            cur = nxt
            i += 1
        if pending is not None:
            yield self._finish(pending)

    def _finish(self, pending):
        hout, hcnt, ev, cap, x_host = pending
        ev.synchronize()
        c = hcnt.tolist()
        B = len(c) - 1
        if c[B] > cap or c[B] < 0 or any(k < 0 for k in c[:B]):  # rare: more candidates than the optimistic capacity -> re-run, blocking
            xd = x_host.to(self.device)
            pred = self.model.detect_records(xd) if self.fused else self.model(xd)[0]
            packed, counts = non_max_suppression_obb(pred, return_packed=True, **self.kw)
            host = packed.cpu()
            return [host[b, :k].clone() for b, k in enumerate(counts)]
        return [hout[b, :k].clone() for b, k in enumerate(c[:B])]
