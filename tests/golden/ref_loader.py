"""Import the UNMODIFIED Python reference: /root/reference in the build container, else the staged copy
oracle/_ref/py/ (oracle/build_ref.py; git-ignored build artefact that gpurun ships to the GPU box).

Used by tests/golden/make_golden.py / make_trace.py to produce the committed fixtures, by the live-forward GPU tests
(tests/test_gpu_live_forward.py) and by bench.py's `forward` leg; nothing under pats_b200/ touches it.  The shims are the ones SURVEY.md Appendix A verified: empty stand-ins for optional
top-level imports the hot path never touches, a kornia-0.5.5 `create_meshgrid`, resnet34
without a download, and `Tensor.cuda` as identity on CPU.  No reference file is edited.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAGED = os.path.join(REPO, "oracle", "_ref", "py")  # oracle/build_ref.py: copy of the reference's models/ + utils/ for the GPU box


def _pick_root() -> str:
    env = os.environ.get("PATS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.exists("/root/reference/models/modules.py"):
        return "/root/reference"
    return STAGED


REF_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "models", "modules.py"))


def _stub(name: str, **attrs):
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
        return mod


def load_reference():
    """Returns a namespace with the reference modules (modules, utils, first/second/third layer, pats)."""
    import torch

    if not reference_available():
        raise FileNotFoundError(REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # compiled reference tensor_resize (setup/library.cpp) must be importable as `tensor_resize`
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from oracle.build_ref import load_ref

    sys.modules["tensor_resize"] = load_ref()

    for name in ("h5py", "imagesize", "pydegensac", "open3d", "plotly", "_plotly_utils"):
        _stub(name)
    _stub("_plotly_utils.basevalidators", ColorscaleValidator=object)
    _stub("numpy.lib.function_base", average=None)

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=0).transpose(1, 2)
        return base.unsqueeze(0).permute(0, 2, 3, 1)

    _stub("kornia")
    _stub("kornia.utils")
    _stub("kornia.utils.grid", create_meshgrid=create_meshgrid)

    import torchvision

    _orig_resnet34 = torchvision.models.resnet34
    if not getattr(_orig_resnet34, "_pats_nodl", False):
        def resnet34(*a, **k):
            k.pop("pretrained", None)
            k["weights"] = None
            return _orig_resnet34(*a, **k)

        resnet34._pats_nodl = True
        torchvision.models.resnet34 = resnet34

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    ns = types.SimpleNamespace()
    ns.modules = importlib.import_module("models.modules")
    ns.utils = importlib.import_module("utils.utils")
    ns.first_layer = importlib.import_module("models.first_layer")
    ns.second_layer = importlib.import_module("models.second_layer")
    ns.third_layer = importlib.import_module("models.third_layer")
    ns.pats = importlib.import_module("models.pats")
    ns.tensor_resize = sys.modules["tensor_resize"]
    return ns
