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
SOURCES = ["api.cu", "sinkhorn.cu", "subdivide.cu", "regroup.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "pats_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", SO, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a CC that nvcc's host pass must not pick up
    subprocess.run(cmd, check=True, env=env)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
