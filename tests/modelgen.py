"""Seeded model state shared by the golden-vector scripts (run on the REFERENCE model) and the tests
(run on the mirror / engine): identical parameters without shipping a checkpoint."""
import torch


def seeded_state(model, seed: int = 0, obj_bias: float = None):
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
        det = model.model[-1]
        with torch.no_grad():
            for mi in det.m:
                b = mi.bias.view(det.na, -1)
                b[:, 4] = obj_bias
                b[:, 5:5 + det.nc] = obj_bias
    return model


def build_mirror(size="n", nc=15, seed=0, obj_bias=None):
    from yolov5_obb_b200 import yolo as Y
    torch.manual_seed(seed)
    m = Y.Model(f"yolov5{size}.yaml", ch=3, nc=nc)
    return seeded_state(m, seed, obj_bias).eval()
