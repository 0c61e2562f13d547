"""Quick per-kernel timing on the GPU box (development aid; bench.py is the contract).
Writes gpurun_out/kernel_times.json."""
import json
import math
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pats_b200 import _lib, modules, tensor_resize, utils  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return {"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "max_ms": ts[-1]}


def torch_ot2(s, ns, iters):
    b, m, n = s.shape
    ms = float(m - 1)
    nsum = ns.sum(2).reshape(b)
    norm = -(ms + nsum).log()
    lnu = torch.cat([ns.reshape(b, -1).log() + norm[:, None], (math.log(ms) + norm)[:, None]], 1)
    lmu = torch.cat([norm[:, None].expand(b, m - 1), (nsum.log() + norm)[:, None]], 1)
    u, v = torch.zeros_like(lmu), torch.zeros_like(lnu)
    for _ in range(iters):
        u = lmu - torch.logsumexp(s + v[:, None, :], 2)
        v = lnu - torch.logsumexp(s + u[:, :, None], 1)
    return s + u[:, :, None] + v[:, None, :] - norm[:, None, None]


res = {"device": torch.cuda.get_device_name(0), "sm_count": lib.pats_sm_count()}
g = torch.Generator().manual_seed(0)
one = torch.tensor(1.0, device=dev)
for name, (b, m, n) in {"L3_4800x65": (4800, 65, 65), "L3_30000x65": (30000, 65, 65), "L3_153600x65": (153600, 65, 65),
                         "L2_300x145": (300, 145, 145), "L2_9600x145": (9600, 145, 145), "L2_148x145": (148, 145, 145)}.items():
    s = (0.1 * torch.randn(b, m, n, generator=g)).to(dev)
    ns = torch.exp((torch.rand(b, 1, n - 1, generator=g) * 2 - 1) * 2.77).to(dev)
    out = torch.empty_like(s)

    def run():
        rc = lib.pats_log_optimal_transport2_f32(s.data_ptr(), one.data_ptr(), ns.data_ptr(), b, m, n, 100, out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0

    if m == 65:
        for mode in (1, 2, 3):
            lib.pats_sinkhorn_disable_w65(mode)
            tv = timeit(run)
            res[f"{name}_w65mode{mode}"] = tv
            print(name, "w65 mode", mode, tv["median_ms"], flush=True)
        lib.pats_sinkhorn_disable_w65(0)
    if m == 145:
        for mode in (1, 2):
            lib.pats_sinkhorn_disable_c145(mode)
            tv = timeit(run)
            res[f"{name}_c145mode{mode}"] = tv
            print(name, "c145 mode", mode, tv["median_ms"], flush=True)
        lib.pats_sinkhorn_disable_c145(0)
    t = timeit(run)
    t["us_per_problem"] = 1e3 * t["median_ms"] / b
    t["alg_GBps"] = b * m * n * 4 * 102 / (t["median_ms"] * 1e-3) / 1e9
    t["ffma_TFLOPs"] = b * m * n * 2 * 2 * 99 / (t["median_ms"] * 1e-3) / 1e12
    res[name] = t
    print(name, t, flush=True)
    if b <= 4800:
        tt = timeit(lambda: torch_ot2(s, ns, 100), reps=3, warm=1)
        res[name + "_torch_cuda"] = tt
        print(name, "torch-cuda", tt, flush=True)
    del s, ns, out
    torch.cuda.empty_cache()

for name, (b, m) in {"L1_1x300": (1, 300), "L1_32x300": (32, 300)}.items():
    s = (0.1 * torch.randn(b, m, m, generator=g)).to(dev)
    ns = torch.exp((torch.rand(b, 1, m, generator=g) * 2 - 1) * 2.77).to(dev)
    for v in (0, 1, 2):
        lib.pats_sinkhorn_cluster_variant(v)
        t = timeit(lambda: modules.log_optimal_transport(s, one, ns, 100), reps=10)
        res[f"{name}_cluster_variant{v}"] = t
        print(name, "cluster variant", v, t, flush=True)
    lib.pats_sinkhorn_cluster_variant(0)

# plans beyond 512 x 512: grid-cooperative streaming kernel against the HBM roofline (SURVEY section 8d:
# B_alg = b * 4 * (M+1)(N+1) * (iters + 2)) and against the one-CTA-per-problem log-domain kernel
try:
    PEAK = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6462.1)
except Exception:
    PEAK = 6462.1
for name, (b, n, iters) in {"BIG_32x1536": (32, 1536, 100), "BIG_1x1024": (1, 1024, 100), "BIG_8x1024": (8, 1024, 100),
                            "BIG_1x4096_it200": (1, 4096, 200)}.items():
    if os.environ.get("KT_ONLY") and not name.startswith(os.environ["KT_ONLY"]):
        continue
    s = (0.1 * torch.randn(b, n, n, generator=g)).to(dev)
    ns = torch.exp((torch.rand(b, 1, n, generator=g) * 2 - 1) * 2.77).to(dev)
    t = timeit(lambda: modules.log_optimal_transport(s, one, ns, iters), reps=5, warm=2)
    balg = b * 4 * (n + 1) * (n + 1) * (iters + 2)
    t["gbs_alg"] = balg / (t["median_ms"] * 1e-3) / 1e9
    t["frac_hbm"] = t["gbs_alg"] / PEAK
    res[name] = t
    print(name, t, flush=True)
    for v in (1,):
        lib.pats_sinkhorn_grid_variant(v)
        tv = timeit(lambda: modules.log_optimal_transport(s, one, ns, iters), reps=5, warm=2)
        lib.pats_sinkhorn_grid_variant(0)
        res[f"{name}_variant{v}"] = tv
        print(name, "variant", v, tv["median_ms"], flush=True)
    if os.environ.get("KT_GENERIC") and b * n <= 8192 * 4:
        lib.pats_sinkhorn_force_generic(1)
        tg = timeit(lambda: modules.log_optimal_transport(s, one, ns, iters), reps=2, warm=1)
        lib.pats_sinkhorn_force_generic(0)
        res[name + "_generic"] = tg
        print(name, "generic", tg, flush=True)
    del s, ns
    torch.cuda.empty_cache()

src = torch.floor(torch.rand(1, 3, 736, 896, generator=g) * 256).to(dev)
rows = []
for k in range(300):
    cy, cx = 128 + 16 + 32 * (k // 20), 128 + 16 + 32 * (k % 20)
    half = int(torch.randint(24, 96, (1,), generator=g))
    rows.append([max(0, cy - half), min(735, cy + half), max(0, cx - half), min(895, cx + half), k])
bound = torch.tensor(rows, dtype=torch.long, device=dev)
tensor_resize.CHECK_BOUNDS = False
res["tensor_resize_300"] = timeit(lambda: tensor_resize.tensor_resize(src, bound))
print("tensor_resize_300", res["tensor_resize_300"], flush=True)
left = torch.randint(0, 256, (1, 480, 640, 3), generator=g).to(torch.uint8).to(dev)
right = torch.randint(0, 256, (1, 480, 640, 3), generator=g).to(torch.uint8).to(dev)
xs = torch.exp((torch.rand(1, 300, generator=g) * 2 - 1) * 0.7).to(dev)
avg = (torch.rand(1, 300, 2, generator=g) * torch.tensor([13.0, 18.0]) + 1.0).to(dev)
nm = torch.zeros(1, 300, dtype=torch.bool, device=dev)
res["Compute_imgs_300"] = timeit(lambda: utils.Compute_imgs(xs, xs, avg, nm, left, right))
print("Compute_imgs_300", res["Compute_imgs_300"], flush=True)
left_pad = torch.randint(0, 256, (1, 3, 544, 704), generator=g).to(torch.uint8).to(dev)
res["origin_extract_300"] = timeit(lambda: utils.origin_extract(left_pad, 32, 20, 15))
print("origin_extract_300", res["origin_extract_300"], flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(REPO, "gpurun_out", "kernel_times.json"), "w"), indent=1)
