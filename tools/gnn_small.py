"""Development aid: the fused attention network at the small batch sizes of a chunked (if_local) forward pass."""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))


def main():
    import live_util as L
    from gnn_check import build_module
    from pats_b200 import gnn as G

    dev = torch.device("cuda:0")
    ref = L.load_reference()
    with torch.no_grad():
        for D, N, layers, Bs in ((128, 65, 10, (3, 10, 71, 500)), (264, 145, 18, (19, 60)), (448, 300, 18, (1,))):
            mod, _ = build_module(ref, 5, D, ["self", "cross"] * (layers // 2), dev)
            for B in Bs:
                x0, x1 = torch.randn(B, D, N, device=dev), torch.randn(B, D, N, device=dev)
                for fn, name in ((lambda: G.attentional_gnn_forward(mod, x0, x1), "ours"), (lambda: mod(x0, x1), "module")):
                    fn(); torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(5):
                        fn()
                    t_issue = (time.perf_counter() - t0) / 5
                    torch.cuda.synchronize()
                    t_all = (time.perf_counter() - t0) / 5
                    print(f"D={D} N={N} B={B:4d} {name:7s} host issue {t_issue * 1e3:7.3f} ms   total {t_all * 1e3:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
