"""A/B of the bulk-copy (TMA engine) staging of the 65 x 65 Sinkhorn kernel against its direct loads (X1 of the round-1 verdict).

    python tools/time_bulk_staging.py > gpurun_out/ab_bulk_staging.json

Times log_optimal_transport2 on [b,65,65] planted problems with pats_sinkhorn_bulk_staging off / on, fixed-point exit off (every
problem runs 100 iterations) and on, and checks the plans are bit-identical.  CUDA events, median of 20, after warm-up.
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from pats_b200 import _lib, modules as M  # noqa: E402

sys.argv = sys.argv[:1]
import bench  # noqa: E402  (the planted workload generator)

dev = torch.device("cuda:0")
lib = _lib.load()
res = {}
for b in (4800, 38400):
    g = torch.Generator().manual_seed(1)
    s, area = bench.planted_scores(torch, g, b, 8, 8, sharp=1.5, noise=0.3, floor=-10.0, dustbin=True, peak=7.0)
    ns = (area.reshape(-1, 1, 1) * torch.exp((torch.rand(b, 1, 64, generator=g) * 2 - 1) * 0.18)).to(dev)
    s = s.to(dev)
    one = torch.tensor(1.0, device=dev)
    ref = None
    for fp in (0, 1):
        lib.pats_sinkhorn_fixed_point_exit(fp)
        for bulk in (0, 1):
            lib.pats_sinkhorn_bulk_staging(bulk)
            out = M.log_optimal_transport2(s, one, ns, 100)
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(out, ref))
            for _ in range(5):
                M.log_optimal_transport2(s, one, ns, 100)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for e0, e1 in evs:
                e0.record()
                M.log_optimal_transport2(s, one, ns, 100)
                e1.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(c) for a, c in evs)
            res[f"b{b}_fpexit{fp}_bulk{bulk}"] = {"median_ms": ts[10], "min_ms": ts[0], "bit_identical_to_direct": same}
            print(f"b={b} fp_exit={fp} bulk={bulk}: median {ts[10]:.4f} ms  min {ts[0]:.4f}  identical {same}", file=sys.stderr, flush=True)
lib.pats_sinkhorn_bulk_staging(1)
lib.pats_sinkhorn_fixed_point_exit(1)
os.write(bench._REAL_STDOUT, (json.dumps(res, indent=1) + "\n").encode())
