"""Triage of the ill-conditioned OT records of tests/golden/trace_*.npz on a GPU (development aid).

For every log_optimal_transport / log_optimal_transport2 record it prints, per record:
  * max|scores| and the f32 spacing there (the quantum the reference's potentials u, v are rounded to),
  * E_ref  = max |reference f32 (CPU, stored) - the same iteration in f64|        (the reference's own rounding error),
  * E_cuda = max |the reference's formulation in f32 on THIS GPU (ATen logsumexp) - stored CPU reference|,
  * ours (default dispatch) and ours (log-domain kernel forced): max |d| and max d / (1e-4 + 2e-6|ref|) against the
    stored reference, max |ours - f64|, the number of problems that took the in-kernel log-domain fallback.

    python tools/triage_trace.py [tag ...] > gpurun_out/triage.json
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import torch  # noqa: E402

import trace_util as T  # noqa: E402


ot_torch = T.ot_reference_torch


def main():
    import numpy as np

    import pats_b200.modules as M
    from pats_b200 import _lib

    lib = _lib.load()
    dev = "cuda:0"
    tags = sys.argv[1:] or ["portrait", "global", "local"]
    rows = []
    for tag in tags:
        z, meta = T.load(tag)
        for c in meta["calls"]:
            name = c["name"]
            if name not in ("log_optimal_transport", "log_optimal_transport2"):
                continue
            args = T.decode(c["args"], z, dev)
            want = T.decode(c["out"], z, dev)
            scores, alpha, ns = args[0], args[1], args[2]
            iters = int(args[3]) if len(args) > 3 else 100
            r64 = ot_torch(name, scores.double(), alpha.double(), ns.double(), iters)
            r32c = ot_torch(name, scores, alpha, ns, iters)
            lim = 1e-4 + 2e-6 * want.double().abs()
            smax = float(scores.abs().max())
            row = {"tag": tag, "seq": c["seq"], "name": name, "shape": list(scores.shape), "max_abs_score": smax,
                   "f32_spacing": float(np.spacing(np.float32(smax))),
                   "E_ref_vs_f64": float((want.double() - r64).abs().max()),
                   "E_torchcuda_vs_ref": float((r32c.double() - want.double()).abs().max()),
                   "E_torchcuda_vs_ref_over_lim": float(((r32c.double() - want.double()).abs() / lim).max())}
            for label, force in (("ours", 0), ("ours_logdomain", 1)):
                lib.pats_sinkhorn_force_generic(force)
                lib.pats_sinkhorn_fallback_count(1)
                fn = M.log_optimal_transport if name == "log_optimal_transport" else M.log_optimal_transport2
                got = fn(scores, alpha, ns, iters)
                torch.cuda.synchronize()
                fb = lib.pats_sinkhorn_fallback_count(0)
                d = (got.double() - want.double()).abs()
                per_problem = (d / lim).flatten(1).max(dim=1).values
                worst = int(torch.argmax((d / lim).flatten()))
                row[label] = {"fallbacks": fb, "max_d": float(d.max()), "max_d_over_lim": float((d / lim).max()),
                              "n_beyond": int((d > lim).sum()), "per_problem_max_over_lim": [round(float(x), 3) for x in per_problem[:16]],
                              "worst_ref": float(want.flatten()[worst]), "worst_got": float(got.flatten()[worst]),
                              "max_vs_f64": float((got.double() - r64).abs().max()),
                              "argmax_rows_equal": bool((got.argmax(2) == want.argmax(2)).all()),
                              "argmax_cols_equal": bool((got.argmax(1) == want.argmax(1)).all())}
            lib.pats_sinkhorn_force_generic(0)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(REPO, "gpurun_out", "triage.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
