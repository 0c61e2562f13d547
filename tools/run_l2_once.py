"""One warm pass + profiled pass of the level-2 composite (2400 planted 145 x 145 problems: solve + area expansion) for ncu."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from pats_b200 import layers as Ly  # noqa: E402

sys.argv = sys.argv[:1]
import bench  # noqa: E402

dev = torch.device("cuda:0")
d = bench.make_inputs(torch, 8, 18027, "planted")
s, ns, sx, sy = (d[k].to(dev) for k in ("l2_scores", "l2_ns", "l2_sx", "l2_sy"))
for _ in range(3):
    Ly.second_layer_match(s, 1.0, ns, sx, sy, 100, True, 12)
torch.cuda.synchronize()
