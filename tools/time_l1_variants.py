"""Times pats_log_optimal_transport_f32 on one 301 x 301 problem (level 1) for every cluster variant
(pats_sinkhorn_cluster_variant); event pair around one wrapper call, minimum of 20."""
import sys, math, torch
sys.path.insert(0, ".")
from pats_b200 import _lib, modules as M
lib = _lib.load()
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
s = (0.1 * torch.randn(1, 300, 300, generator=g)).to(dev)
ns = torch.exp((torch.rand(1, 1, 300, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
alpha = torch.tensor(1.0, device=dev)
ref = None
for v in (0, 1, 2, 3, 4):  # 0 auto, 1 = 4 x 512, 2 = 8 x 256 one-hop, 3 = 8 x 256, 4 = 10 x 256
    lib.pats_sinkhorn_cluster_variant(v)
    try:
        out = M.log_optimal_transport(s, alpha, ns, 100)
        torch.cuda.synchronize()
    except Exception as e:
        print(v, "FAILED", e); continue
    if ref is None: ref = out.clone()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = M.log_optimal_transport(s, alpha, ns, 100); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    print(f"variant {v}: min {min(ts):.1f} us median {sorted(ts)[10]:.1f} us maxdiff {float((out - ref).abs().max()):.2e}")
lib.pats_sinkhorn_cluster_variant(0)
