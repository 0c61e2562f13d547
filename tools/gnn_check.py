"""Development aid: accuracy and timing of the fused attention network (csrc/gnn.cu) against the reference module on the same GPU.

    python tools/gnn_check.py            # writes gpurun_out/gnn_check.json
Accuracy: the golden cases of tests/golden/gnn.npz (truth: oracle/gnn.py in float64); timing: the shapes of one 640x480 pair."""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))


def build_module(ref, seed, D, names, dev):
    from oracle import gnn as O

    params = O.seeded_params(seed, len(names), D)
    gnn = ref.modules.AttentionalGNN(D, names).eval()
    sd = {}
    for l, p in enumerate(params):
        for k, v in p.items():
            sd[f"layers.{l}.{k}"] = torch.from_numpy(v)
        sd[f"layers.{l}.mlp.1.num_batches_tracked"] = torch.tensor(0)
    gnn.load_state_dict(sd, strict=True)
    return gnn.to(dev), params


def main():
    import live_util as L
    from make_gnn_golden import CASES, inputs
    from oracle import gnn as O
    from pats_b200 import gnn as G

    dev = torch.device("cuda:0")
    ref = L.load_reference()
    out = {"accuracy": {}, "timing": {}}
    with torch.no_grad():
        for name in ("tiny", "l3", "l2"):
            seed, B, D, N, names = CASES[name]
            mod, params = build_module(ref, seed, D, names, dev)
            d0, d1 = inputs(seed, B, D, N)
            t0, t1 = O.attentional_gnn(params, names, d0, d1)
            scale = float(max(np.abs(t0).max(), np.abs(t1).max()))
            x0, x1 = torch.from_numpy(d0).to(dev), torch.from_numpy(d1).to(dev)
            r0, r1 = mod(x0, x1)
            rec = {"scale": scale, "aten_cuda_err": float(max(np.abs(r0.cpu().numpy() - t0).max(), np.abs(r1.cpu().numpy() - t1).max()))}
            from pats_b200 import _lib
            for variant in (0, 1):
                _lib.load().pats_gnn_gemm_variant(variant)
                for passes in (3, 1):
                    G.set_precision(passes)
                    o0, o1 = G.attentional_gnn_forward(mod, x0, x1)
                    torch.cuda.synchronize()
                    rec[f"gemm{variant}_{passes}x_err"] = float(max(np.abs(o0.cpu().numpy() - t0).max(), np.abs(o1.cpu().numpy() - t1).max()))
                    if variant == 0 and passes == 3:
                        keep = o0.clone()
                    if variant == 1 and passes == 3:
                        rec["variants_bit_identical_3x"] = bool(torch.equal(keep, o0))
            _lib.load().pats_gnn_gemm_variant(0)
            G.set_precision(3)
            out["accuracy"][name] = rec
            print(name, rec, flush=True)
        for name, B, D, N, names in (("l3_K2800", 2800, 128, 65, ["self", "cross"] * 5), ("l2_P40", 40, 264, 145, ["self", "cross"] * 9),
                                     ("l2_P300", 300, 264, 145, ["self", "cross"] * 9)):
            mod, _ = build_module(ref, 5, D, names, dev)
            g = torch.Generator().manual_seed(1)
            x0, x1 = torch.randn(B, D, N, generator=g).to(dev), torch.randn(B, D, N, generator=g).to(dev)
            rec = {}

            def timed(fn, reps=3):
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            rec["reference_module_ms"] = timed(lambda: mod(x0, x1))
            from pats_b200 import _lib
            for variant in (0, 1):
                _lib.load().pats_gnn_gemm_variant(variant)
                for passes in (3, 1):
                    G.set_precision(passes)
                    rec[f"gemm{variant}_{passes}x_ms"] = timed(lambda: G.attentional_gnn_forward(mod, x0, x1))
                    if passes == 3:
                        for mb in (128, 1024):
                            packed, cross, _, heads, _ = G.pack_module(mod)
                            rec[f"gemm{variant}_{passes}x_ws{mb}_ms"] = timed(lambda: G.attentional_gnn(packed, cross, heads, x0, x1, workspace_mb=mb))
            _lib.load().pats_gnn_gemm_variant(0)
            G.set_precision(3)
            a0, _ = G.attentional_gnn_forward(mod, x0, x1)
            b0, _ = mod(x0, x1)
            rec["max_abs_diff_vs_module"] = float((a0 - b0).abs().max())
            rec["scale"] = float(b0.abs().max())
            out["timing"][name] = rec
            print(name, rec, flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(REPO, "gpurun_out", "gnn_check.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
