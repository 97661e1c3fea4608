"""Shared by make_loss_golden.py and the loss tests."""
import torch

CASES = {"a": dict(B=2, imgsz=128, nt=24, seed=1), "b": dict(B=3, imgsz=256, nt=60, seed=2),
         "empty": dict(B=2, imgsz=128, nt=0, seed=3), "smooth": dict(B=2, imgsz=128, nt=16, seed=4, ls=0.1)}
STRIDES = [8.0, 16.0, 32.0]
_ANCH = torch.tensor([[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]).float()
ANCHORS_GRID = _ANCH.view(3, 3, 2) / torch.tensor(STRIDES).view(3, 1, 1)  # Detect.anchors after yolo.py:124


def hyp_from_golden(G, name):
    v = G[f"{name}/hyp"]
    return dict(box=float(v[0]), obj=float(v[1]), cls=float(v[2]), theta=float(v[3]), anchor_t=float(v[4]),
                label_smoothing=float(v[5]), cls_pw=1.0, obj_pw=1.0, theta_pw=1.0, fl_gamma=0.0)


class FakeModel:
    """What ComputeLoss reads from a model: .hyp and .model[-1] (Detect)."""

    def __init__(self, hyp, nc=15):
        from yolov5_obb_b200 import yolo as Y
        self.hyp = hyp
        det = Y.Detect(nc=nc, anchors=_ANCH.view(3, -1).tolist(), ch=(8, 8, 8))
        det.stride = torch.tensor(STRIDES)
        det.anchors /= det.stride.view(-1, 1, 1)
        self.model = [det]
