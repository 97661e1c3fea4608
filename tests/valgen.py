"""Seeded inputs for the validation-matching parity (val.py:209-250): per image a set of ground-truth rboxes (network-input
pixels), detections = jittered copies of some of them (several per object, some with the wrong class) + false positives, sorted
by descending confidence like non_max_suppression_obb's output, and the dataloader's letterbox `shapes` entries."""
import numpy as np

PI = 3.141592
SHAPES = [((800, 1000), ((1.024, 1.024), (0.0, 102.4))),      # (h_raw, w_raw), ((gain, gain), (pad_x, pad_y))
          ((1024, 1024), ((1.0, 1.0), (0.0, 0.0))),
          ((600, 450), ((1.7066667079925537, 1.7066667079925537), (128.0, 0.0)))]


def synth_val_batch(seed=0, max_det=300, nc=15, imgsz=1024):
    rng = np.random.default_rng(seed)
    B = len(SHAPES)
    dets = np.zeros((B, max_det, 7), np.float32)
    counts = np.zeros(B, np.int64)
    targets = []
    for b in range(B):
        ng = [40, 0, 120][b]                                    # image 1 has no labels (val.py:244-245 branch)
        c = rng.uniform(60, imgsz - 60, (ng, 2))
        l = np.exp(rng.uniform(np.log(10), np.log(200), ng))
        s = l * rng.uniform(0.2, 1.0, ng)
        th = (rng.integers(0, 180, ng) - 90) / 180 * PI
        cls = rng.integers(0, nc, ng)
        for i in range(ng):
            targets.append([b, cls[i], c[i, 0], c[i, 1], l[i], s[i], th[i]])
        rows = []
        for i in range(ng):
            for _ in range(rng.integers(0, 4)):                 # 0-3 detections per object
                q = [0.005, 0.05, 0.15][rng.integers(0, 3)]        # near-exact, good, sloppy localisation
                jit = rng.normal(0, q, 2) * s[i]
                k = cls[i] if rng.random() > 0.15 else rng.integers(0, nc)
                rows.append([c[i, 0] + jit[0], c[i, 1] + jit[1], l[i] * (1 + rng.uniform(-2, 2) * q), s[i] * (1 + rng.uniform(-2, 2) * q),
                             th[i] + rng.normal(0, q), rng.uniform(0.05, 1.0), k])
        for _ in range([30, 50, 40][b]):                        # false positives
            ll = np.exp(rng.uniform(np.log(10), np.log(200)))
            rows.append([rng.uniform(0, imgsz), rng.uniform(0, imgsz), ll, ll * rng.uniform(0.2, 1.0),
                         (rng.integers(0, 180) - 90) / 180 * PI, rng.uniform(0.05, 1.0), rng.integers(0, nc)])
        rows = np.asarray(rows, np.float32)
        rows = rows[np.argsort(-rows[:, 5], kind="stable")][:max_det]
        dets[b, :len(rows)] = rows
        counts[b] = len(rows)
    return dets, counts, np.asarray(targets, np.float32), SHAPES
