"""Seeded synthetic rotated boxes shaped like SURVEY §8(d): long edge log-uniform 8–300 px,
aspect U(0.15,1), θ on the 180-bin CSL grid with the reference's truncated π, unique scores."""
import numpy as np

PI = 3.141592  # utils/rboxs_utils.py:5


def rboxes(n, span, seed, n_classes=15, class_offset=True, theta_grid=True):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, span, (n, 2))
    l = np.exp(rng.uniform(np.log(8), np.log(300), n))
    s = l * rng.uniform(0.15, 1.0, n)
    if theta_grid:
        th = (rng.integers(0, 180, n) - 90) / 180 * PI
    else:
        th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    d = np.stack([c[:, 0], c[:, 1], l, s, th], 1).astype(np.float32)
    cls = rng.integers(0, n_classes, n)
    if class_offset:  # utils/general.py:849-851
        d[:, :2] = d[:, :2] + (cls[:, None].astype(np.float32) * np.float32(4096))
    scores = ((rng.permutation(n) + 1).astype(np.float32)) / np.float32(n)
    return d, scores, cls


def degenerate_pairs():
    """Pairs that stress the thresholds in box_iou_rotated_utils.h (parallel edges, shared corners,
    identical boxes, containment, tiny and zero-area boxes)."""
    A, B = [], []

    def add(a, b):
        A.append(a)
        B.append(b)

    for th in (0.0, PI / 2 * 0.5, -PI / 2, 0.3, 1.0):
        add([10, 10, 20, 10, th], [10, 10, 20, 10, th])            # identical
        add([10, 10, 20, 10, th], [10, 10, 10, 20, th])            # crossed
        add([10, 10, 20, 10, th], [30, 10, 20, 10, th])            # shared edge along x (if th=0)
        add([10, 10, 20, 10, th], [10, 10, 5, 2, th + 0.1])        # contained
        add([10, 10, 20, 10, th], [10.5, 10.25, 20, 10, th])       # slightly shifted
        add([1000, 2000, 300, 40, th], [1001, 2001, 280, 45, -th])
    add([0, 0, 1e-8, 1e-8, 0], [0, 0, 1, 1, 0])                    # area < 1e-14
    add([0, 0, 0, 0, 0], [0, 0, 0, 0, 0])
    add([5, 5, 2, 2, 0], [7, 7, 2, 2, 0])                          # touching at a corner
    add([5, 5, 2, 2, 0], [7, 5, 2, 2, 0])                          # touching along an edge
    add([61440.5, 61441.25, 33.3, 12.1, 0.2], [61442.5, 61440.0, 30.0, 14.0, -0.4])  # class-offset magnitudes
    # the same rectangle written two ways (w,h swapped, theta +- 90 deg): every edge pair is parallel or
    # coincident — SURVEY 8(c)'s known-answer boxes 2 and 3 are this case
    q = np.pi / 4
    add([100, 100, 141.4, 141.4, -q], [100, 100, 141.4, 141.4, q])
    add([100, 100, 50, 20, 0.3], [100, 100, 20, 50, 0.3 - np.pi / 2])
    add([100, 100, 50, 20, 0.0], [100, 100, 20, 50, np.pi / 2])
    add([4196, 4196.5, 64, 16, -PI / 2], [4196, 4196.5, 16, 64, 0.0])
    return np.asarray(A, np.float32), np.asarray(B, np.float32)
