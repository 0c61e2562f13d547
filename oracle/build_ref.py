"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference native op into oracle/_ref/.

The one native component of zju3dv/pats is `setup/library.cpp` (pybind11 module
``tensor_resize``, reference setup/library.cpp:47-66,92-95; built upstream by
setup/setup.py:114-115 as a plain torch CppExtension).  This recipe compiles that
file *where it lies* under /root/reference with g++ directly (no reference build
system, no source copied into this repo) and writes ``oracle/_ref/tensor_resize*.so``.

oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot;
there is no /root/reference on that box, so only the prebuilt .so is used there.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may load the result.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PATS_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
SRC = os.path.join(REF_ROOT, "setup", "library.cpp")


def ref_so_path() -> str:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(OUT_DIR, "tensor_resize" + suffix)


def build_ref(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference library.cpp -> oracle/_ref/tensor_resize.*.so.

    Returns the path, or None when /root/reference is absent (GPU box) and no
    prebuilt file exists.
    """
    out = ref_so_path()
    if not os.path.exists(SRC):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [
        "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
        "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=tensor_resize",
        f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        *inc, SRC, "-o", out,
        f"-L{libdir}", f"-Wl,-rpath,{libdir}",
        "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
    ]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


PY_DIR = os.path.join(OUT_DIR, "py")


def stage_reference_python(force: bool = False) -> str | None:
    """Stage the reference's Python hot-path callers for the GPU box: models/*.py and utils/*.py are copied VERBATIM from
    /root/reference into oracle/_ref/py/ (a git-ignored build output, like the compiled op above: it never enters the
    history, and gpurun ships it).  The live-forward tests and bench.py's `forward` leg import the UNMODIFIED
    models/pats.py from there when /root/reference does not exist.  Returns the directory, or None if neither exists."""
    import shutil

    src_ok = os.path.exists(os.path.join(REF_ROOT, "models", "pats.py"))
    if not src_ok:
        return PY_DIR if os.path.exists(os.path.join(PY_DIR, "models", "pats.py")) else None
    for pkg in ("models", "utils"):
        dst = os.path.join(PY_DIR, pkg)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(REF_ROOT, pkg))):
            if not f.endswith(".py"):
                continue
            a, b = os.path.join(REF_ROOT, pkg, f), os.path.join(dst, f)
            if force or not os.path.exists(b) or os.path.getmtime(b) < os.path.getmtime(a):
                shutil.copyfile(a, b)
    return PY_DIR


def load_ref():
    """Import the compiled reference module (needs torch imported first)."""
    import importlib.util

    import torch  # noqa: F401  (registers the libtorch symbols the .so needs)

    path = build_ref()
    if path is None or not os.path.exists(path):
        raise FileNotFoundError("oracle/_ref/tensor_resize*.so not built and /root/reference absent")
    spec = importlib.util.spec_from_file_location("tensor_resize", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build_ref(force="--force" in sys.argv, verbose=True)
    print("built:" if p else "unavailable:", p)
    print("staged python:", stage_reference_python(force="--force" in sys.argv))
