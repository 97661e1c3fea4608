"""Seeded model state shared by the golden-vector scripts (run on the REFERENCE model) and the tests
(run on the mirror / engine): identical parameters without shipping a checkpoint."""
import torch


def seeded_state(model, seed: int = 0, obj_bias: float = None, cls_bias: float = None, det_gain: float = 1.0):
    """Give BatchNorm non-trivial affine + running statistics (the reference's __init__ leaves them at
    the values of a zero-image probe) and optionally raise the Detect obj/cls biases so NMS has work."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            with torch.no_grad():
                m.weight.copy_(torch.rand(n, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
                m.num_batches_tracked.zero_()
    if obj_bias is not None:
        # random-init Detect logits barely vary across locations (std ~0.04), so a threshold passes all anchors or
        # none; widening the obj/cls rows gives the NMS stage a DOTA-like load of a few thousand candidates per tile
        det = model.model[-1]
        with torch.no_grad():
            for mi in det.m:
                w = mi.weight.view(det.na, det.no, -1)
                w[:, 4:5 + det.nc] *= det_gain
                b = mi.bias.view(det.na, -1)
                b[:, 4] = obj_bias
                b[:, 5:5 + det.nc] = obj_bias if cls_bias is None else cls_bias
    return model


def build_mirror(size="n", nc=15, seed=0, obj_bias=None, cls_bias=None, det_gain=1.0):
    from yolov5_obb_b200 import yolo as Y
    torch.manual_seed(seed)
    m = Y.Model(f"yolov5{size}.yaml", ch=3, nc=nc)
    return seeded_state(m, seed, obj_bias, cls_bias, det_gain).eval()


def calibrated_bench_model(size="s", nc=15, seed=0, obj_frac=0.015, cls_frac=0.05, conf=0.25):
    """Seeded random-init model made 'alive' for the benchmark, identically on every arm (CPU fp32, deterministic):
    (1) BatchNorm running statistics := batch statistics of one synthetic calibration tile (what training would
    have produced; with the raw seeded statistics the signal decays to a constant and a confidence threshold
    passes every anchor or none), (2) per (level, anchor) the Detect obj row is rescaled so that ~obj_frac of
    the anchors exceed `conf` and each class row so that ~cls_frac exceed 0.5: a few thousand spatially
    irregular NMS candidates per tile, like DOTA val (test.txt: NMS is ~1/4 of the per-image time)."""
    import math
    from oracle import model_ref
    from tests.tilegen import synth_tiles
    m = build_mirror(size, nc=nc, seed=seed)
    x = synth_tiles(1, 1024, seed=999).float() / 255
    bns = [b for b in m.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    for b in bns:
        b.momentum = 1.0
    with torch.no_grad():
        raw = model_ref.forward(m, x, training=True)  # updates running stats in place, returns raw logits per level
    for b in bns:
        b.momentum = 0.03
    det = m.model[-1]
    logit_conf = math.log(conf / (1 - conf))
    with torch.no_grad():
        for l, mi in enumerate(det.m):
            w = mi.weight.view(det.na, det.no, -1)
            bias = mi.bias.view(det.na, det.no)
            r = raw[l][0]  # [na, H, W, no]
            for a in range(det.na):
                for c in range(4, 5 + nc):
                    v = r[a, :, :, c].flatten().double()
                    g = 2.0 / max(v.std().item(), 1e-6)
                    frac, target = (obj_frac, logit_conf) if c == 4 else (cls_frac, 0.0)
                    q = torch.quantile(v, 1.0 - frac).item()
                    w[a, c] *= g
                    bias[a, c] = g * bias[a, c] + (target - g * q)
    return m.eval()
