"""Development aid for ncu: one fused attention-network call at a pair's level-2 shape (P x 264 x 145, 18 layers) or level-3 shape.
    ncu --set full --clock-control none --import-source on -k regex:gnn_gemm -s 6 -c 3 -o gpurun_out/gnn python tools/run_gnn_once.py l2 300"""
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

    level = sys.argv[1] if len(sys.argv) > 1 else "l2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    D, N = (264, 145) if level == "l2" else (128, 65)
    dev = torch.device("cuda:0")
    ref = L.load_reference()
    with torch.no_grad():
        mod, _ = build_module(ref, 5, D, ["self", "cross"] * (layers // 2), dev)
        g = torch.Generator().manual_seed(1)
        x0, x1 = torch.randn(B, D, N, generator=g).to(dev), torch.randn(B, D, N, generator=g).to(dev)
        G.attentional_gnn_forward(mod, x0, x1)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
