"""Where the time of one PATS.forward goes once the hot path is installed.

Runs the unmodified reference `models/pats.py` (oracle/_ref/py on the GPU box) with `install(fused=True)` on synthetic 640x480
pairs and reports (i) wall time per sub-module, measured with synchronising pre/post forward hooks (so the figures are
exclusive of overlap: they answer "what is left", not "how fast could it be"), (ii) the top CUDA kernels and the CPU-side
time of a torch.profiler trace of one pair, (iii) the number of kernel launches per pair.  Development aid; writes
gpurun_out/profile_forward.json.
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    import torch
    import live_util as L
    import pats_b200.install as inst

    dev = torch.device("cuda:0")
    ref = L.load_reference()
    cfg = L.config(if_local="global" not in sys.argv[1:], merge_new=True, if_outdoor=True)
    pairs = [L.synthetic_pair((480, 640), seed=L.SEED + i) for i in range(4)]
    out = {}
    with torch.no_grad():
        model = L.build_model(ref, cfg, device=dev)
        attention = "attention" in sys.argv[1:]
        inst.install(fused=True, attention=attention)
        out["attention"] = attention
        try:
            def run(p):
                r = model({"image0": p[0].to(dev), "image1": p[1].to(dev)})
                return int(r["matches_l"].cpu().shape[0])

            run(pairs[0]); run(pairs[1])
            torch.cuda.synchronize()
            t0 = time.perf_counter(); run(pairs[2]); torch.cuda.synchronize()
            out["s_per_pair_free_running"] = time.perf_counter() - t0

            # (i) synchronising hooks on the modules of interest
            acc, stack, handles = {}, [], []

            def pre(name):
                def h(mod, inp):
                    torch.cuda.synchronize(); stack.append((name, time.perf_counter()))
                return h

            def post(name):
                def h(mod, inp, o):
                    torch.cuda.synchronize(); n, t = stack.pop(); a = acc.setdefault(n, [0.0, 0]); a[0] += time.perf_counter() - t; a[1] += 1
                return h

            names = {}
            for top in ("first_layer", "second_layer", "third_layer"):
                m = getattr(model, top)
                names[top] = m
                for cn, c in m.named_children():
                    names[top + "." + cn] = c
            for n, m in names.items():
                handles.append(m.register_forward_pre_hook(pre(n)))
                handles.append(m.register_forward_hook(post(n)))
            torch.cuda.synchronize(); t0 = time.perf_counter(); run(pairs[3]); torch.cuda.synchronize()
            out["s_per_pair_synchronised"] = time.perf_counter() - t0
            for h in handles:
                h.remove()
            out["modules_s"] = {k: {"s": round(v[0], 5), "calls": v[1]} for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])}

            # (ii) profiler trace of one pair
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                run(pairs[2]); torch.cuda.synchronize()
            ev = prof.key_averages()
            kern = [(e.key, e.device_time_total, e.count) for e in ev if e.device_type == torch.autograd.DeviceType.CUDA]
            kern.sort(key=lambda x: -x[1])
            out["cuda_kernel_total_ms"] = sum(k[1] for k in kern) / 1e3
            out["cuda_kernel_launches"] = sum(k[2] for k in kern)
            out["top_kernels"] = [{"name": k[0][:110], "ms": round(k[1] / 1e3, 3), "n": k[2]} for k in kern[:40]]
            cpu = [(e.key, e.self_cpu_time_total, e.count) for e in ev if e.device_type == torch.autograd.DeviceType.CPU]
            cpu.sort(key=lambda x: -x[1])
            out["cpu_self_total_ms"] = sum(k[1] for k in cpu) / 1e3
            out["top_cpu_ops"] = [{"name": k[0][:80], "ms": round(k[1] / 1e3, 3), "n": k[2]} for k in cpu[:25]]
        finally:
            inst.uninstall()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", ("profile_forward_global" if "global" in sys.argv[1:] else "profile_forward") + ("_attention.json" if out["attention"] else ".json")), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("s_per_pair_free_running", "s_per_pair_synchronised", "cuda_kernel_total_ms", "cuda_kernel_launches", "cpu_self_total_ms")}))


if __name__ == "__main__":
    main()
