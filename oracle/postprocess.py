"""CPU restatement of /root/reference/utils/general.py:772-862 (non_max_suppression_obb) —
TEST INFRASTRUCTURE.  torch CPU ops for the filtering, the pinned C++ oracle for the rotated NMS
(mode 1 = reference CUDA rule `>`, mode 0 = reference CPU rule `>=`).  Pinned by
tests/golden/postprocess_golden.npz (outputs of the reference function itself)."""
import numpy as np
import torch

from . import obb_nms as _oracle_obb_nms

PI = 3.141592  # utils/rboxs_utils.py:5 (general.py imports it)


def non_max_suppression_obb(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                            multi_label=False, max_det=1500, nms_mode=1):
    prediction = prediction.detach().float().cpu()
    nc = prediction.shape[2] - 5 - 180
    xc = prediction[..., 4] > conf_thres
    class_index = nc + 5
    max_wh, max_nms = 4096, 30000
    multi_label &= nc > 1
    output = [torch.zeros((0, 7))] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]].clone()
        if not x.shape[0]:
            continue
        x[:, 5:class_index] *= x[:, 4:5]
        _, theta_pred = torch.max(x[:, class_index:], 1, keepdim=True)
        theta_pred = (theta_pred - 90) / 180 * PI
        if multi_label:
            i, j = (x[:, 5:class_index] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((x[i, :4], theta_pred[i], x[i, j + 5, None], j[:, None].float()), 1)
        else:
            conf, j = x[:, 5:class_index].max(1, keepdim=True)
            x = torch.cat((x[:, :4], theta_pred, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 6:7] == torch.tensor(classes)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 5].argsort(descending=True, stable=True)[:max_nms]]
        c = x[:, 6:7] * (0 if agnostic else max_wh)
        rboxes = x[:, :5].clone()
        rboxes[:, :2] = rboxes[:, :2] + c
        keep = _oracle_obb_nms(rboxes.numpy(), x[:, 5].numpy(), float(np.float32(iou_thres)), mode=nms_mode)
        keep = torch.from_numpy(keep)
        if keep.shape[0] > max_det:
            keep = keep[:max_det]
        output[xi] = x[keep]
    return output
