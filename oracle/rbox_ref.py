"""numpy/torch-CPU restatement of /root/reference/utils/rboxs_utils.py:9-26,106-165 and
utils/general.py:636-650 — TEST INFRASTRUCTURE.  Pinned by tests/golden/rbox_golden.npz."""
import numpy as np

PI = 3.141592


def gaussian_label_cpu(label, num_class, u=0, sig=4.0):
    x = np.arange(-num_class / 2, num_class / 2)
    y_sig = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    index = int(num_class / 2 - label)
    return np.concatenate([y_sig[index:], y_sig[:index]], axis=0)


def rbox2poly(obboxes):
    center, w, h, theta = np.split(obboxes.astype(np.float32), (2, 3, 4), axis=-1)
    Cos, Sin = np.cos(theta), np.sin(theta)
    vector1 = np.concatenate([w / 2 * Cos, -w / 2 * Sin], axis=-1)
    vector2 = np.concatenate([-h / 2 * Sin, -h / 2 * Cos], axis=-1)
    point1 = center + vector1 + vector2
    point2 = center + vector1 - vector2
    point3 = center - vector1 - vector2
    point4 = center - vector1 + vector2
    return np.concatenate([point1, point2, point3, point4], axis=-1).reshape(*obboxes.shape[:-1], 8)


def poly2hbb(polys):
    x, y = polys[:, 0::2], polys[:, 1::2]
    x_max, x_min, y_max, y_min = x.max(1), x.min(1), y.max(1), y.min(1)
    return np.stack([(x_max + x_min) / 2.0, (y_max + y_min) / 2.0, x_max - x_min, y_max - y_min], 1)


def scale_polys(img1_shape, polys, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    polys = polys.copy()
    polys[:, [0, 2, 4, 6]] -= pad[0]
    polys[:, [1, 3, 5, 7]] -= pad[1]
    polys[:, :8] /= gain
    return polys
