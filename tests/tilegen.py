"""Seeded synthetic DOTA-shaped uint8 tiles: low-frequency background + ~23 oriented bright rectangles per tile
(n ~ Poisson(23.4), long edge log-uniform 8..300 px, aspect U(0.15, 1)).  Pure torch, deterministic."""
import math

import torch


def synth_tiles(batch: int, size: int = 1024, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + seed)
    out = torch.empty((batch, 3, size, size), dtype=torch.uint8)
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
    for b in range(batch):
        low = torch.rand((1, 3, size // 32, size // 32), generator=g)
        img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)[0] * 0.5 + 0.15
        img = img + (torch.rand((3, size, size), generator=g) - 0.5) * 0.08
        n = int(torch.poisson(torch.tensor(23.4), generator=g).clamp(1, 200).item())
        for _ in range(n):
            cx, cy = (torch.rand(2, generator=g) * size).tolist()
            l = math.exp(math.log(8) + torch.rand(1, generator=g).item() * (math.log(300) - math.log(8)))
            s = l * (0.15 + 0.85 * torch.rand(1, generator=g).item())
            th = (torch.rand(1, generator=g).item() - 0.5) * math.pi
            col = torch.rand(3, generator=g) * 0.6 + 0.4
            r = int(l / 2 + 2)
            x0, x1, y0, y1 = max(int(cx) - r, 0), min(int(cx) + r + 1, size), max(int(cy) - r, 0), min(int(cy) + r + 1, size)
            if x0 >= x1 or y0 >= y1:
                continue
            dx, dy = xx[y0:y1, x0:x1] - cx, yy[y0:y1, x0:x1] - cy
            u = dx * math.cos(th) + dy * math.sin(th)
            v = -dx * math.sin(th) + dy * math.cos(th)
            msk = (u.abs() <= l / 2) & (v.abs() <= s / 2)
            patch = img[:, y0:y1, x0:x1]
            patch[:, msk] = col[:, None].expand(3, int(msk.sum()))
        out[b] = (img.clamp(0, 1) * 255).to(torch.uint8)
    return out
