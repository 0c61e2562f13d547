"""One warm launch + one profiled launch of the level-3 Sinkhorn (38 400 planted 65 x 65 problems = 8 pairs, 100 iterations) -- the
target of the ncu captures under profiles/ (r02_w65x2_*).  PATS_L3_FP_EXIT=0 / PATS_L3_BULK=0 select the A/B variants."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from pats_b200 import _lib, modules as M  # noqa: E402

sys.argv = sys.argv[:1]
import bench  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
lib.pats_sinkhorn_fixed_point_exit(int(os.environ.get("PATS_L3_FP_EXIT", "1")))
lib.pats_sinkhorn_bulk_staging(int(os.environ.get("PATS_L3_BULK", "1")))
b = int(os.environ.get("PATS_L3_B", "38400"))
g = torch.Generator().manual_seed(1)
s, area = bench.planted_scores(torch, g, b, 8, 8, sharp=1.5, noise=0.3, floor=-10.0, dustbin=True, peak=7.0)
ns = (area.reshape(-1, 1, 1) * torch.exp((torch.rand(b, 1, 64, generator=g) * 2 - 1) * 0.18)).to(dev)
s = s.to(dev)
one = torch.tensor(1.0, device=dev)
for _ in range(3):
    M.log_optimal_transport2(s, one, ns, 100)
torch.cuda.synchronize()
