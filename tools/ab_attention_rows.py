"""A/B of the level-2 attention tiling (query rows per warp x warps per CTA): python tools/ab_attention_rows.py [out.json]

Variants of pats_gnn_attention_variant: 0 = 8 rows x 20 warps (shipped), 2 = 16 rows x 10 warps, 3 = 12 rows x 13 warps.  Prints the device
time of one attention-network call (CUDA events, best of 5) and of its attention launches (CUPTI), and whether the outputs agree bit for bit."""
import json
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
    from pats_b200 import gnn as G, _lib
    from torch.profiler import profile, ProfilerActivity

    lib = _lib.load()
    dev = torch.device("cuda:0")
    ref = L.load_reference()
    out = {}
    with torch.no_grad():
        for name, B in (("l2_P300", 300), ("l2_P56", 56)):
            mod, _ = build_module(ref, 5, 264, ["self", "cross"] * 9, dev)
            g = torch.Generator().manual_seed(1)
            x0, x1 = torch.randn(B, 264, 145, generator=g).to(dev), torch.randn(B, 264, 145, generator=g).to(dev)
            base = None
            for av in (0, 2, 3, 0):
                lib.pats_gnn_attention_variant(av)
                y = G.attentional_gnn_forward(mod, x0, x1)
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    y = G.attentional_gnn_forward(mod, x0, x1)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    G.attentional_gnn_forward(mod, x0, x1)
                    torch.cuda.synchronize()
                att = [e for e in prof.key_averages() if "gnn_attention" in e.key]
                att_ms = sum(e.device_time_total for e in att) / 1e3
                if base is None:
                    base = [t.clone() for t in y]
                same = all(torch.equal(a, b) for a, b in zip(base, y))
                out.setdefault(name, []).append({"variant": av, "network_ms": round(best, 3), "attention_ms": round(att_ms, 3),
                                                 "attention_launches": sum(e.count for e in att), "bit_identical_to_variant_0": bool(same)})
                print(name, out[name][-1], flush=True)
    lib.pats_gnn_attention_variant(0)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
