"""Seeded quadrilaterals for the devkit polygon-IoU / tile-merge NMS parity (rotated rectangles as DOTA results are
written by ResultMerge: 8 coordinates per detection, any orientation and winding, plus general convex quads and
degenerate / identical / disjoint pairs)."""
import numpy as np


def _rect(c, l, s, th):
    cs, sn = np.cos(th), np.sin(th)
    pts = np.array([[-l / 2, -s / 2], [l / 2, -s / 2], [l / 2, s / 2], [-l / 2, s / 2]]) @ np.array([[cs, sn], [-sn, cs]]) + c
    return pts.reshape(8)


def quad_pairs(seed=0, n=3000):
    rng = np.random.default_rng(seed)
    P, Q = [], []
    for i in range(n):
        c = rng.uniform(0, 4000, 2)
        p = _rect(c, rng.uniform(5, 300), rng.uniform(3, 120), rng.uniform(-np.pi, np.pi))
        kind = i % 6
        if kind == 0:      # far apart
            q = _rect(c + 2000, rng.uniform(5, 300), rng.uniform(3, 120), rng.uniform(-np.pi, np.pi))
        elif kind == 1:    # identical
            q = p.copy()
        elif kind == 2:    # same box, other winding / start corner
            q = np.roll(p.reshape(4, 2)[::-1], rng.integers(0, 4), 0).reshape(8)
        elif kind == 3:    # general convex quad around the same centre
            ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
            r = rng.uniform(5, 150, 4)
            q = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1).reshape(8)
        else:              # overlapping rotated rectangles
            q = _rect(c + rng.normal(0, 20, 2), rng.uniform(5, 300), rng.uniform(3, 120), rng.uniform(-np.pi, np.pi))
        if i % 4 == 0:
            p, q = np.round(p), np.round(q)   # result files carry 1 decimal at most; integers create collinear edges
        P.append(p)
        Q.append(q)
    return np.stack(P).astype(np.float64), np.stack(Q).astype(np.float64)


def merge_dets(n, seed=0):
    """[n, 9]: clustered rotated boxes with unique scores, as several overlapping tiles report the same objects."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0, 3000, (max(4, n // 6), 2))
    rows = []
    for i in range(n):
        c = centres[rng.integers(0, len(centres))] + rng.normal(0, 4, 2)
        l = rng.uniform(20, 200)
        rows.append(np.concatenate([_rect(c, l, l * rng.uniform(0.2, 1.0), rng.uniform(-np.pi, np.pi)), [0.0]]))
    D = np.stack(rows)
    D[:, 8] = rng.permutation(n) / n * 0.9 + 0.05
    return D
