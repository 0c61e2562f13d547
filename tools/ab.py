"""A/B harness for the Sinkhorn kernels (development aid; bench.py is the contract).

    python tools/ab.py build  name1:-DFLAG_A=1  name2:"-DFLAG_B=2 -DFLAG_C"   # here (CPU box): tools/ab/lib_<name>.so
    python tools/ab.py run [L3] [L2] [L1] [BIG]                               # on the GPU box: times every built variant

A variant is sinkhorn.cu + sinkhorn_grid.cu + api.cu compiled with extra -D flags; `base` (no flags) is always built.
Every variant is checked against `base` (max |diff| of the plans) before it is timed.
"""
import ctypes as C
import glob
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "pats_b200", "csrc")
OUT = os.path.join(REPO, "tools", "ab")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def build(specs):
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, "lib_*.so")):
        os.remove(f)
    env = dict(os.environ)
    env.pop("CC", None)
    procs = []
    for spec in ["base:"] + specs:
        name, _, flags = spec.partition(":")
        csrc = CSRC
        if "@" in name:  # name@gitrev: sources of that revision (e.g. old@HEAD)
            name, _, rev = name.partition("@")
            root = os.path.join("/tmp", "ab_src_" + name)
            csrc = os.path.join(root, "pats_b200", "csrc")
            os.makedirs(csrc, exist_ok=True)
            os.makedirs(os.path.join(root, "include"), exist_ok=True)
            for rel in ["include/pats_b200.h"] + ["pats_b200/csrc/" + f for f in ("api.cu", "sinkhorn.cu", "sinkhorn_grid.cu", "common.cuh", "sinkhorn_common.cuh")]:
                open(os.path.join(root, rel), "w").write(subprocess.run(["git", "-C", REPO, "show", f"{rev}:{rel}"], capture_output=True, text=True, check=True).stdout)
        so = os.path.join(OUT, f"lib_{name}.so")
        cmd = ["/usr/local/cuda/bin/nvcc", *FLAGS, *flags.split(), "-o", so] + [os.path.join(csrc, s) for s in ("api.cu", "sinkhorn.cu", "sinkhorn_grid.cu", "sinkhorn_c145.cu") if os.path.exists(os.path.join(csrc, s))]
        procs.append((name, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        errs = [l for l in out.splitlines() if "error" in l.lower()]
        print(name, "OK" if p.returncode == 0 else "FAILED", *errs[:5], sep="\n  " if errs else " ")


def run(which):
    import torch

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    one = torch.tensor(1.0, device=dev)
    cases = {}
    if not which or "L3" in which:
        cases["L3_4800x65"] = ("ot2", 4800, 65)
        cases["L3_30000x65"] = ("ot2", 30000, 65)
    if not which or "L2" in which:
        cases["L2_300x145"] = ("ot2", 300, 145)
        cases["L2_2960x145"] = ("ot2", 2960, 145)
    if not which or "L1" in which:
        cases["L1_1x300"] = ("ot", 1, 300)
    if "BIG" in which:
        cases["BIG_32x1536"] = ("ot", 32, 1536)
        cases["BIG_1x1024"] = ("ot", 1, 1024)
    libs = {}
    for so in sorted(glob.glob(os.path.join(OUT, "lib_*.so"))):
        lib = C.CDLL(so)
        for fn in ("pats_log_optimal_transport_f32", "pats_log_optimal_transport2_f32"):
            getattr(lib, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.pats_last_error.restype = C.c_char_p
        libs[os.path.basename(so)[4:-3]] = lib
    res = {}
    for cname, (kind, b, n) in cases.items():
        if kind == "ot2":
            s = (0.1 * torch.randn(b, n, n, generator=g)).to(dev)
            ns = torch.exp((torch.rand(b, 1, n - 1, generator=g) * 2 - 1) * 2.77).to(dev)
            out = torch.empty_like(s)
            m = n
        else:
            s = (0.1 * torch.randn(b, n, n, generator=g)).to(dev)
            ns = torch.exp((torch.rand(b, 1, n, generator=g) * 2 - 1) * 2.77).to(dev)
            out = torch.empty(b, n + 1, n + 1, device=dev)
            m = n
        ref = None
        for vname, lib in libs.items():
            fn = lib.pats_log_optimal_transport2_f32 if kind == "ot2" else lib.pats_log_optimal_transport_f32

            def call():
                rc = fn(s.data_ptr(), one.data_ptr(), ns.data_ptr(), b, m, m, 100, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                assert rc == 0, lib.pats_last_error()

            out.zero_()
            call()
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
                diff = 0.0
            else:
                diff = float((out - ref).abs().max())
            for _ in range(3):
                call()
            reps = 20 if b * n * n < 2e8 else 5
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, e in evs:
                a.record()
                call()
                e.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(e) for a, e in evs)
            res[f"{cname}/{vname}"] = {"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "maxdiff_vs_base": diff}
            print(f"{cname:14s} {vname:24s} median {ts[len(ts) // 2]:8.4f} ms  min {ts[0]:8.4f}  maxdiff {diff:.2e}", flush=True)
        del s, ns, out, ref
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(REPO, "gpurun_out", "ab.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
