"""Seeded polygon sets for poly2rbox parity: DOTA-like rotated rectangles (labelTxt rows are 4 corner points of an
oriented box, rounded to integers), general convex quads, point orders clockwise / counter-clockwise / rotated start,
squares, and degenerate quads (repeated and collinear points)."""
import numpy as np


def p2r_polys(seed=0, n_rect=600, n_quad=300):
    rng = np.random.default_rng(seed)
    polys = []
    # rotated rectangles, any orientation, corner order rotated / reversed at random
    for i in range(n_rect):
        cx, cy = rng.uniform(0, 1024, 2)
        l = np.exp(rng.uniform(np.log(6), np.log(400)))
        s = l * rng.uniform(0.1, 1.0)
        th = rng.uniform(-np.pi, np.pi)
        c, sn = np.cos(th), np.sin(th)
        pts = np.array([[-l / 2, -s / 2], [l / 2, -s / 2], [l / 2, s / 2], [-l / 2, s / 2]])
        pts = pts @ np.array([[c, sn], [-sn, c]]) + (cx, cy)
        if i % 3 == 0:
            pts = np.round(pts)           # DOTA annotations are integer pixels
        pts = np.roll(pts, rng.integers(0, 4), 0)
        if rng.random() < 0.5:
            pts = pts[::-1]
        polys.append(pts.reshape(8))
    # general convex quads (not rectangles)
    for i in range(n_quad):
        c = rng.uniform(100, 900, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        r = rng.uniform(10, 200, 4)
        pts = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
        polys.append(pts.reshape(8))
    # axis-aligned, squares, degenerate
    polys += [np.array(v, float) for v in (
        [0, 0, 100, 0, 100, 50, 0, 50], [0, 0, 50, 0, 50, 100, 0, 100], [10, 10, 60, 10, 60, 60, 10, 60],
        [1707, 1539, 1683, 1523, 1689, 1513, 1713, 1529], [0, 0, 10, 10, 20, 20, 30, 30], [5, 5, 5, 5, 5, 5, 5, 5],
        [0, 0, 40, 0, 40, 0, 0, 30], [100, 100, 140, 140, 100, 180, 60, 140], [0, 0, 100, 1, 100, 51, 0, 50])]
    return np.stack(polys).astype(np.float64)
