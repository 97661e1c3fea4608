"""Seeded synthetic loss inputs: Detect training outputs and DOTA-shaped targets with CSL rows
(targets[:, 7:] follow utils/rboxs_utils.py:9-26 gaussian_label_cpu, restated in numpy)."""
import numpy as np
import torch

PI = 3.141592


def gaussian_label(angle, num_class=180, sig=2.0):
    """utils/rboxs_utils.py:9-26 (u=0): peak at bin 90 - trunc(90 - angle)."""
    x = np.arange(-num_class / 2, num_class / 2)
    y = np.exp(-(x ** 2) / (2 * sig ** 2))
    index = int(num_class / 2 - angle)
    return np.concatenate([y[index:], y[:index]], axis=0)


def synth_targets(B, nt, imgsz, nc=15, seed=0):
    rng = np.random.default_rng(seed)
    t = np.zeros((nt, 187), np.float32)
    t[:, 0] = rng.integers(0, B, nt)
    t[:, 1] = rng.integers(0, nc, nt)
    t[:, 2:4] = rng.uniform(0, imgsz, (nt, 2))
    l = np.exp(rng.uniform(np.log(8), np.log(min(300, imgsz)), nt))
    t[:, 4] = l
    t[:, 5] = l * rng.uniform(0.15, 1.0, nt)
    k = rng.integers(0, 180, nt)
    theta = (k - 90) / 180 * PI
    t[:, 6] = theta
    for i in range(nt):
        angle = theta[i] * 180 / PI + 90
        t[i, 7:] = gaussian_label(angle)
    # duplicates: two targets in the same cell exercise the last-writer-wins tobj rule
    if nt >= 4:
        t[1, :] = t[0, :]
        t[1, 4:6] *= 1.1
        t[3, 2:4] = t[2, 2:4] + 0.5
    return t


def synth_preds(B, imgsz, no=200, seed=0, na=3, scale=1.5):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, na, imgsz // s, imgsz // s, no, generator=g) * scale for s in (8, 16, 32)]
