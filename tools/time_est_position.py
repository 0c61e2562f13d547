"""Times the level-2 area expansion (pats_est_position_f32, b = 300, 12 x 12 grid) on diffuse plans (0.1*randn scores: every
box grows in all 8 iterations) and on peaked plans (a planted warp: boxes stop after a few iterations -> fixed-point exit)."""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from pats_b200 import layers as Ly  # noqa: E402
from pats_b200 import modules as M  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(18027)
b, n = 300, 144
ns = torch.exp((torch.rand(b, 1, n, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
sx = ns.reshape(b, n).sqrt().contiguous()
ys, xs = torch.meshgrid(torch.arange(12.0), torch.arange(12.0), indexing="ij")
src = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
out = {}
for tag, sharp in (("diffuse", 0.0), ("peaked", 6.0), ("sharp", 80.0)):
    sc = 0.1 * torch.randn(b, n + 1, n + 1, generator=g)
    if sharp:
        t = torch.randn(b, 1, 2, generator=g) * 1.5
        d2 = ((src[None, :, None, :] + t[:, :, None, :] - src[None, None, :, :]) ** 2).sum(-1)
        sc[:, :n, :n] += -sharp * d2 * 0.1
    Z = M.log_optimal_transport2(sc.to(dev), 1.0, ns, 100)
    for _ in range(3):
        Ly.est_position(Z, sx, sx, 12, 12, 8, 1e-3)
    ts = []
    for _ in range(20):  # one event pair per call and the minimum: the kernel, not the wrapper's host time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = Ly.est_position(Z, sx, sx, 12, 12, 8, 1e-3, return_extra=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    bound = r[-1]
    out[tag] = {"us_min": min(ts), "us_median": sorted(ts)[len(ts) // 2],
                "mean_box_cells": float(((bound[..., 1] - bound[..., 0] + 1) * (bound[..., 3] - bound[..., 2] + 1)).float().mean())}
print(json.dumps(out))
