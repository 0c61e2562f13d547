"""The tcgen05 correlation kernel against the reference's three-op formulation (einsum, / sqrt(d), * 0.1) on the shapes of one
640x480 pair and of a batch of eight: CUDA events, median of 20.   python tools/time_correlation.py > gpurun_out/ab_correlation.json"""
import json
import math
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from pats_b200 import layers as Ly  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
res = {}


def timed(fn):
    for _ in range(5):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[10]


for name, b, d, n in (("L3_4800x128x65", 4800, 128, 65), ("L3_38400x128x65", 38400, 128, 65), ("L2_300x264x145", 300, 264, 145),
                      ("L2_2400x264x145", 2400, 264, 145), ("L1_1x448x300", 1, 448, 300)):
    g = torch.Generator().manual_seed(1)
    d0 = torch.randn(b, d, n, generator=g).to(dev)
    d1 = torch.randn(b, d, n, generator=g).to(dev)
    sc = 0.1 / math.sqrt(d)
    t_ref = timed(lambda: 0.1 * (torch.einsum('bdn,bdm->bnm', d0, d1) / d ** .5))
    t_gemm = timed(lambda: torch.einsum('bdn,bdm->bnm', d0, d1))
    t_ours = timed(lambda: Ly.correlation(d0, d1, sc))
    err = float((Ly.correlation(d0, d1, sc) - 0.1 * (torch.einsum('bdn,bdm->bnm', d0, d1) / d ** .5)).abs().max())
    res[name] = {"reference_three_ops_ms": t_ref, "einsum_alone_ms": t_gemm, "tcgen05_ms": t_ours, "max_abs_diff": err}
    print(name, res[name], file=sys.stderr, flush=True)
    del d0, d1
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
