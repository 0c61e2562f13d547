"""Times the composite level-2 / level-3 calls against their parts (development aid)."""
import math, os, sys, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pats_b200 import _lib, layers as Ly, modules as M
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)

def timeit(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]

res = {}
for b in (296, 300, 148, 152):
    n = 144
    s = (0.1 * torch.randn(b, n + 1, n + 1, generator=g)).to(dev)
    sx = torch.exp((torch.rand(b, n, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    sy = torch.exp((torch.rand(b, n, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    ns = (sx * sy).reshape(b, 1, n)
    Z = M.log_optimal_transport2(s, 1.0, ns, 100)
    r = {"ot2": timeit(lambda: M.log_optimal_transport2(s, 1.0, ns, 100)), "est2": timeit(lambda: Ly.est_position(Z, sx, sy, 12, 12, 8, 1e-3))}
    for h in (1, 0):
        lib.pats_plan_handover(h)
        r[f"composite_handover{h}"] = timeit(lambda: Ly.second_layer_match(s, 1.0, ns, sx, sy, 100, True, 12))
    lib.pats_plan_handover(1)
    res[f"L2_b{b}"] = r
    print(f"L2 b={b}", {k: round(v, 4) for k, v in r.items()}, flush=True)
for K in (4736, 4800, 9472):
    s = (0.1 * torch.randn(K, 65, 65, generator=g)).to(dev)
    ns = torch.exp((torch.rand(K, 1, 64, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    sxy = (ns.reshape(K, 64) + 1e-8).sqrt()
    p_s = (torch.randint(0, 24, (K, 2), generator=g) * 4).to(dev)
    p_t = (torch.randint(0, 25, (K, 2), generator=g) * 4).to(dev)
    Z = M.log_optimal_transport2(s, 1.0, ns, 100)
    r = {"ot3": timeit(lambda: M.log_optimal_transport2(s, 1.0, ns, 100)), "third": timeit(lambda: Ly.third_result_from_log(Z, sxy, sxy, p_s, p_t))}
    for h in (1, 0):
        lib.pats_plan_handover(h)
        r[f"composite_handover{h}"] = timeit(lambda: Ly.third_layer_match(s, 1.0, ns, sxy, sxy, p_s, p_t, 100))
    lib.pats_plan_handover(1)
    res[f"L3_K{K}"] = r
    print(f"L3 K={K}", {k: round(v, 4) for k, v in r.items()}, flush=True)
json.dump(res, open(os.path.join(REPO, "gpurun_out", "composites.json"), "w"), indent=1)
