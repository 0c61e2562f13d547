"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck are 10-100x slower than a plain
run, so sizes and iteration counts are tiny; numerics are checked elsewhere).

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Covers the kernels with inter-thread / inter-CTA protocols: the 10-CTA and 8-CTA cluster kernels (DSMEM st.async + mbarrier
exchange), the 145x145 and 65x65 register kernels (shared-memory row exchange, pair barriers), the composite calls (plan
hand-over by per-problem flags + programmatic dependent launch), the grid-cooperative streaming kernel, area expansion (column
maxima through shared / global atomics, last-CTA reset) and the match assembly (ballot + scan).
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from pats_b200 import _lib, layers as Ly, modules as M, utils as U  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
lib = _lib.load()
IT = int(os.environ.get("PATS_SAN_ITERS", "6"))
one = torch.tensor(1.0, device=dev)


def areas(*shape):
    return torch.exp((torch.rand(*shape, generator=g) * 2 - 1) * 1.0).to(dev)


# level 1: 301 x 301, both cluster shapes
s1 = (0.5 * torch.randn(1, 300, 300, generator=g)).to(dev)
ns1 = areas(1, 1, 300)
for variant in (0, 3):
    lib.pats_sinkhorn_cluster_variant(variant)
    Z1 = M.log_optimal_transport(s1, one, ns1, IT)
lib.pats_sinkhorn_cluster_variant(0)
Ly.est_position(Z1, ns1, ns1, 15, 20, 15, 1e-5)
# level 2 composite (hand-over) and plain
s2 = (0.5 * torch.randn(5, 145, 145, generator=g)).to(dev)
sx, sy = areas(5, 144), areas(5, 144)
out2 = Ly.second_layer_match(s2, 1.0, (sx * sy).reshape(5, 1, 144), sx, sy, IT, True, 12)
M.log_optimal_transport2(s2, one, (sx * sy).reshape(5, 1, 144), IT)
nm_L1 = torch.ones(1, 300, dtype=torch.bool, device=dev)
nm_L1[0, 40:45] = False
Ly.merge_patches_new(None, 5, out2[1].clone(), [480, 640], nm_L1, out2[5].clone(), torch.zeros(1, 300, 16, 9, dtype=torch.float64, device=dev))
# level 3 composite and plain, an odd count (a pair CTA with one idle half)
s3 = (0.5 * torch.randn(19, 65, 65, generator=g)).to(dev)
ns3 = areas(19, 1, 64)
sxy = (ns3.reshape(19, 64) + 1e-8).sqrt()
ps = (torch.randint(0, 24, (19, 2), generator=g) * 4).to(dev)
Ly.third_layer_match(s3, 1.0, ns3, sxy, sxy, ps, ps, IT)
M.log_optimal_transport2(s3, one, ns3, IT)
# ill-conditioned problems: the in-kernel log-domain fallback of the register kernels
M.log_optimal_transport2(s3 * 400.0, one, ns3, IT)
M.log_optimal_transport2(s2 * 400.0, one, (sx * sy).reshape(5, 1, 144), IT)
# tcgen05 correlation (TMEM allocation, mbarrier commit, shared-memory operand staging) on the three shapes, feeding a solve
for (bb, dd, nn) in ((5, 128, 65), (2, 264, 145), (1, 448, 300)):
    c0 = torch.randn(bb, dd, nn, generator=g).to(dev)
    c1 = torch.randn(bb, dd, nn, generator=g).to(dev)
    zc = Ly.correlation(c0, c1, 0.1 / dd ** 0.5)
    if nn == 65:
        M.log_optimal_transport2(zc, one, areas(bb, 1, 64), IT)
# the bulk-staged 65 x 65 kernel with more problems than resident CTAs would need is covered above (19 problems, one CTA each);
# an unaligned view takes the direct kernel
M.log_optimal_transport2(s3[1:], one, ns3[1:], IT)
# streaming (grid-cooperative) kernel and the generic log-domain kernel
sb = (0.3 * torch.randn(2, 600, 600, generator=g)).to(dev)
M.log_optimal_transport(sb, one, areas(2, 1, 600), 3)
# 4096 core columns: a row split over four warps (named barriers within the quad), and the one-warp-per-row kernel it replaced
sq = (0.3 * torch.randn(1, 40, 4097, generator=g)).to(dev)
M.log_optimal_transport2(sq, one, areas(1, 1, 4096), 3)
lib.pats_sinkhorn_grid_variant(2)
M.log_optimal_transport2(sq, one, areas(1, 1, 4096), 3)
lib.pats_sinkhorn_grid_variant(0)
lib.pats_sinkhorn_force_generic(1)
M.log_optimal_transport(s1, one, ns1, 2)
lib.pats_sinkhorn_force_generic(0)
# subdivision + match assembly
left = torch.randint(0, 256, (1, 480, 640, 3), generator=g, dtype=torch.uint8).to(dev)
xs = areas(1, 300)
avg = (torch.rand(1, 300, 2, generator=g) * torch.tensor([13.0, 18.0]) + 1.0).to(dev)
nm = (torch.rand(1, 300, generator=g) < 0.9).to(dev)
U.Compute_imgs(xs, xs, avg, nm, left, left, width=20, height=15)
nm1 = (torch.rand(7, 2304, generator=g) < 0.8).to(dev)
nm0 = torch.ones(1, 300, dtype=torch.bool, device=dev)
nm0[0, 10:17] = False
sc0 = torch.ones(1, 300, 2, device=dev)
U.get_result(1, [nm0, nm1], [avg, (torch.rand(7, 2304, 2, generator=g) * 48).to(dev)], [sc0, torch.ones(7, 2304, 2, device=dev)], [[32, 15, 20], [2, 48, 48]], None)
torch.cuda.synchronize()
print("sanitize_smoke: done")
