"""Level-2 solve (2400 / 300 planted 145 x 145 problems) with the fixed-point exit off / on: CUDA events, median of 20."""
import json
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
res = {}
for pairs in (1, 8):
    for kind in ("planted", "diffuse"):
        d = bench.make_inputs(torch, pairs, 18027, kind)
        s, ns = d["l2_scores"].to(dev), d["l2_ns"].to(dev)
        one = torch.tensor(1.0, device=dev)
        ref = None
        for fp in (0, 1):
            lib.pats_sinkhorn_fixed_point_exit(fp)
            lib.pats_sinkhorn_iterations_skipped(1)
            out = M.log_optimal_transport2(s, one, ns, 100)
            torch.cuda.synchronize()
            skipped = lib.pats_sinkhorn_iterations_skipped(1)
            ref = out.clone() if ref is None else ref
            for _ in range(3):
                M.log_optimal_transport2(s, one, ns, 100)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for a, c in evs:
                a.record()
                M.log_optimal_transport2(s, one, ns, 100)
                c.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(c) for a, c in evs)
            res[f"b{s.shape[0]}_{kind}_exit{fp}"] = {"median_ms": ts[10], "identical": bool(torch.equal(out, ref)),
                                                      "iters_executed_mean": 100 - skipped / s.shape[0]}
            print(f"b={s.shape[0]} {kind} exit={fp}: {ts[10]:.4f} ms identical {torch.equal(out, ref)} iters {100 - skipped / s.shape[0]:.1f}", file=sys.stderr, flush=True)
lib.pats_sinkhorn_fixed_point_exit(1)
os.write(bench._REAL_STDOUT, (json.dumps(res, indent=1) + "\n").encode())
