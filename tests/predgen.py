"""Seeded synthetic Detect outputs [B, A, nc+185] (decoded boxes in pixels, obj/cls/theta probabilities)
with clustered boxes so that NMS has real work.  Deterministic from the seed, so golden fixtures only need
to store the reference function's OUTPUT."""
import numpy as np


def synth_pred(B=2, A=3000, nc=15, seed=0, frac_obj=0.3, span=1024.0, clusters=None, n_tiny=0):
    rng = np.random.default_rng(seed)
    no = nc + 185
    p = np.zeros((B, A, no), np.float32)
    for b in range(B):
        K = clusters if clusters else max(A // 8, 1)
        oc = rng.uniform(0, span, (K, 2))
        ol = np.exp(rng.uniform(np.log(8), np.log(300), K))
        os_ = ol * rng.uniform(0.15, 1.0, K)
        ot = rng.integers(0, 180, K)
        ocls = rng.integers(0, nc, K)
        k = rng.integers(0, K, A)
        jitter = rng.normal(0, 1, (A, 2)) * (os_[k, None] * 0.15)
        p[b, :, 0:2] = oc[k] + jitter
        p[b, :, 2] = ol[k] * rng.uniform(0.85, 1.15, A)
        p[b, :, 3] = os_[k] * rng.uniform(0.85, 1.15, A)
        hi = rng.random(A) < frac_obj
        p[b, :, 4] = np.where(hi, rng.uniform(0.3, 1.0, A), rng.uniform(0.0, 0.2, A))
        cls = rng.uniform(0.0, 0.35, (A, nc))
        cls[np.arange(A), ocls[k]] = rng.uniform(0.5, 1.0, A)
        second = rng.integers(0, nc, A)
        boost = rng.random(A) < 0.2
        cls[np.arange(A)[boost], second[boost]] = rng.uniform(0.4, 0.9, boost.sum())
        p[b, :, 5:5 + nc] = cls
        th = rng.uniform(0.0, 0.1, (A, 180))
        tb = (ot[k] + rng.integers(-1, 2, A)) % 180
        th[np.arange(A), tb] = rng.uniform(0.5, 1.0, A)
        p[b, :, 5 + nc:] = th
    # a few degenerate rows: tiny boxes (dropped by obb_nms), exact duplicates
    p[:, 0, 2] = 1e-4
    p[:, 1, :] = p[:, 2, :]
    if n_tiny:  # degenerate boxes with high scores: they rank inside the top-max_nms and are dropped only afterwards
        r2 = np.random.default_rng(seed + 1000)
        for b in range(B):
            rows = r2.choice(A, n_tiny, replace=False)
            p[b, rows, 2 + r2.integers(0, 2, n_tiny)] = 1e-4
            p[b, rows, 4] = r2.uniform(0.8, 1.0, n_tiny)
    return p
