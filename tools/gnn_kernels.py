"""Development aid: per-kernel device time of one fused attention-network call (torch.profiler / CUPTI).  python tools/gnn_kernels.py [3|1]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))


def main():
    import live_util as L
    from gnn_check import build_module
    from pats_b200 import gnn as G
    from torch.profiler import profile, ProfilerActivity

    G.set_precision(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
    from pats_b200 import _lib
    _lib.load().pats_gnn_attention_variant(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    _lib.load().pats_gnn_gemm_variant(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    dev = torch.device("cuda:0")
    ref = L.load_reference()
    with torch.no_grad():
        for name, B, D, N, names in (("l3_K2800", 2800, 128, 65, ["self", "cross"] * 5), ("l2_P300", 300, 264, 145, ["self", "cross"] * 9),
                                     ("l2_P40", 40, 264, 145, ["self", "cross"] * 9)):
            mod, _ = build_module(ref, 5, D, names, dev)
            g = torch.Generator().manual_seed(1)
            x0, x1 = torch.randn(B, D, N, generator=g).to(dev), torch.randn(B, D, N, generator=g).to(dev)
            G.attentional_gnn_forward(mod, x0, x1)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                G.attentional_gnn_forward(mod, x0, x1)
                torch.cuda.synchronize()
            print(name)
            for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
                print(f"   {e.key[:90]:90s} n={e.count:5d} total={e.device_time_total / 1e3:8.3f} ms  avg={e.device_time_total / max(e.count, 1):8.1f} us")


if __name__ == "__main__":
    main()
