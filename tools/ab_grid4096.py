"""A/B of the streaming Sinkhorn kernel at 4096 core columns (BASELINE.json's stress size): four warps per row, exponentials kept
(default) against one warp per row with the exponentials recomputed (pats_sinkhorn_grid_variant(2)).

    python tools/ab_grid4096.py [out.json]

Parity of BOTH against a torch float32 logsumexp restatement on the same GPU (the check of tests/test_gpu_ot.py::
test_large_plan_grid_kernel_vs_torch), in all three modes; timing of one problem at 200 iterations (CUDA events, median of 5)."""
import json
import math
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def lse_sinkhorn(Z, lmu, lnu, iters):
    u, v = torch.zeros_like(lmu), torch.zeros_like(lnu)
    for _ in range(iters):
        u = lmu - torch.logsumexp(Z + v[:, None, :], 2)
        v = lnu - torch.logsumexp(Z + u[:, :, None], 1)
    return Z + u[:, :, None] + v[:, None, :]


def main():
    from pats_b200 import _lib, modules as M

    lib = _lib.load()
    dev = torch.device("cuda:0")
    out = {"library": os.environ.get("PATS_B200_LIB", "default"), "parity": [], "timing": []}
    N = 4096
    if not os.environ.get("AB_TIMING_ONLY"):
        parity(lib, M, dev, N, out)
    timing(lib, M, dev, N, out)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


def parity(lib, M, dev, N, out):
    # ---- log_optimal_transport (dustbin synthesised): b = 2, 40 iterations ----
    g = torch.Generator().manual_seed(8000 + N)
    b, iters = 2, 40
    s = (0.1 * torch.randn(b, N, N, generator=g)).to(dev)
    ns = torch.exp((torch.rand(b, 1, N, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    alpha = torch.tensor(1.0, device=dev)
    Z = torch.cat([torch.cat([s, alpha.expand(b, N, 1)], 2), alpha.expand(b, 1, N + 1)], 1)
    nsum = ns.sum(2).reshape(b)
    norm = -(N + nsum).log()
    lnu = torch.cat([ns.reshape(b, N).log() + norm[:, None], (math.log(N) + norm)[:, None]], 1)
    lmu = torch.cat([norm[:, None].expand(b, N), (nsum.log() + norm)[:, None]], 1)
    ref = lse_sinkhorn(Z, lmu, lnu, iters) - norm[:, None, None]
    res = {}
    for v in (0, 2):
        lib.pats_sinkhorn_grid_variant(v)
        o = M.log_optimal_transport(s, alpha, ns, iters)
        torch.cuda.synchronize()
        res[v] = o
        out["parity"].append({"mode": "ot", "b": b, "shape": [N + 1, N + 1], "iters": iters, "variant": v, "max_abs_diff_vs_torch_lse": float((o - ref).abs().max()),
                              "argmax_equal": bool(torch.equal(o.argmax(2), ref.argmax(2))), "finite": bool(torch.isfinite(o).all())})
        print(out["parity"][-1], flush=True)
    # ---- log_optimal_transport2: dustbin in memory, row stride 4097 (unaligned rows), M small and M odd ----
    for m, iters in ((37, 12), (601, 9)):
        g = torch.Generator().manual_seed(8100 + m)
        s2 = (0.1 * torch.randn(2, m, N + 1, generator=g)).to(dev)
        ns2 = torch.exp((torch.rand(2, 1, N, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
        nsum = ns2.sum(2).reshape(2)
        ms = float(m - 1)
        norm = -(ms + nsum).log()
        lnu = torch.cat([ns2.reshape(2, N).log() + norm[:, None], (math.log(ms) + norm)[:, None]], 1)
        lmu = torch.cat([norm[:, None].expand(2, m - 1), (nsum.log() + norm)[:, None]], 1)
        ref = lse_sinkhorn(s2, lmu, lnu, iters) - norm[:, None, None]
        for v in (0, 2):
            lib.pats_sinkhorn_grid_variant(v)
            o = M.log_optimal_transport2(s2, 1.0, ns2, iters)
            torch.cuda.synchronize()
            out["parity"].append({"mode": "ot2", "b": 2, "shape": [m, N + 1], "iters": iters, "variant": v, "max_abs_diff_vs_torch_lse": float((o - ref).abs().max()),
                                  "argmax_equal": bool(torch.equal(o.argmax(2), ref.argmax(2))), "finite": bool(torch.isfinite(o).all())})
            print(out["parity"][-1], flush=True)
    # ---- log_sinkhorn_iterations (raw marginals), 0 / 1 / 2 / 7 iterations ----
    g = torch.Generator().manual_seed(8200)
    s3 = (0.5 * torch.randn(1, 530, N + 1, generator=g)).to(dev)
    lmu = torch.log_softmax(torch.randn(1, 530, generator=g), 1).to(dev)
    lnu = torch.log_softmax(torch.randn(1, N + 1, generator=g), 1).to(dev)
    for iters in (0, 1, 2, 7):
        ref = lse_sinkhorn(s3, lmu, lnu, iters)
        for v in (0, 2):
            lib.pats_sinkhorn_grid_variant(v)
            o = M.log_sinkhorn_iterations(s3, lmu, lnu, iters)
            torch.cuda.synchronize()
            out["parity"].append({"mode": "raw", "b": 1, "shape": [530, N + 1], "iters": iters, "variant": v, "max_abs_diff_vs_torch_lse": float((o - ref).abs().max()),
                                  "argmax_equal": bool(torch.equal(o.argmax(2), ref.argmax(2))), "finite": bool(torch.isfinite(o).all())})
            print(out["parity"][-1], flush=True)
    lib.pats_sinkhorn_grid_variant(0)


def timing(lib, M, dev, N, out):
    # ---- timing: bench.py's stress leg (one problem, 200 iterations) and b = 2 ----
    for b in ((1,) if os.environ.get("AB_ONLY4096") else (1, 2)):
        g2 = torch.Generator().manual_seed(1234 + N)
        sc = (0.1 * torch.randn(b, N, N, generator=g2)).to(dev)
        nss = torch.exp((torch.rand(b, 1, N, generator=g2) * 2 - 1) * math.log(16.0)).to(dev)
        one = torch.tensor(1.0, device=dev)
        for v in ((0, 2, 0, 2) if not os.environ.get("AB_TIMING_ONLY") else (0, 2)):
            lib.pats_sinkhorn_grid_variant(v)
            for _ in range(2):
                M.log_optimal_transport(sc, one, nss, 200)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for e0, e1 in ev:
                e0.record()
                M.log_optimal_transport(sc, one, nss, 200)
                e1.record()
            torch.cuda.synchronize()
            ms = sorted(x.elapsed_time(y) for x, y in ev)[2]
            nbytes = 4 * b * (N + 1) * (N + 1) * 202
            out["timing"].append({"b": b, "iters": 200, "variant": v, "ms": round(ms, 4), "us_per_iteration": round(ms * 1e3 / 200, 2), "algorithmic_gbs": round(nbytes / ms / 1e6, 1)})
            print(out["timing"][-1], flush=True)
    lib.pats_sinkhorn_grid_variant(0)
    # ---- the other streaming shapes (same kernel family, shared exchange code): BASELINE configs[2], the 1025 x 1025 level-1 plan ----
    for b, n, iters in (() if os.environ.get("AB_ONLY4096") else ((32, 1536, 100), (1, 1024, 100), (8, 1024, 100))):
        g2 = torch.Generator().manual_seed(77 + n + b)
        sc = (0.1 * torch.randn(b, n, n, generator=g2)).to(dev)
        nss = torch.exp((torch.rand(b, 1, n, generator=g2) * 2 - 1) * math.log(16.0)).to(dev)
        one = torch.tensor(1.0, device=dev)
        for _ in range(3):
            M.log_optimal_transport(sc, one, nss, iters)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
        for e0, e1 in ev:
            e0.record()
            M.log_optimal_transport(sc, one, nss, iters)
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in ev)[3]
        nbytes = 4 * b * (n + 1) * (n + 1) * (iters + 2)
        out["timing"].append({"b": b, "N": n, "iters": iters, "ms": round(ms, 4), "us_per_iteration": round(ms * 1e3 / iters, 2), "algorithmic_gbs": round(nbytes / ms / 1e6, 1)})
        print(out["timing"][-1], flush=True)
        del sc, nss


if __name__ == "__main__":
    main()
