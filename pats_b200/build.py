"""Build recipe for libpats_b200.so (hand-written CUDA for sm_100a, C ABI in include/pats_b200.h).

    python -m pats_b200.build [--force]

Plain nvcc; the library does not link against torch.  The .so is built in-tree (git-ignored,
shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SO = os.path.join(HERE, "libpats_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
# source -> extra flags.  regroup.cu is compiled without FMA contraction: its f32 expressions must round
# exactly like the C oracle's (integer results hang off them).
SOURCES = {"api.cu": [], "sinkhorn.cu": [], "sinkhorn_grid.cu": [], "subdivide.cu": [], "regroup.cu": ["-fmad=false"], "gather.cu": [], "correlation.cu": [], "gnn.cu": []}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def headers():
    """Every header a translation unit may include: an edit to any of them rebuilds every object (the TUs share struct layouts)."""
    import glob

    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(INCLUDE, "*.h")))


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a CC that nvcc's host pass must not pick up
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps_t = max(os.path.getmtime(d) for d in headers())
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), deps_t):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *SOURCES[os.path.basename(src)], *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, env=env)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *objs]
    if verbose:
        print(" ".join(link))
    subprocess.run(link, check=True, env=env)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
