"""One launch of the streaming (grid-cooperative) Sinkhorn kernel on BASELINE.json configs[2] -- the target of the ncu capture
whose DRAM bytes go into profiles/roofline_traffic.json (roofline_streaming.traffic)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pats_b200 import modules as M
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(18027)
s = (0.1 * torch.randn(32, 1536, 1536, generator=g)).to(dev)
ns = torch.exp((torch.rand(32, 1, 1536, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
one = torch.tensor(1.0, device=dev)
for _ in range(2):
    M.log_optimal_transport(s, one, ns, 100)
torch.cuda.synchronize()
