"""Make the unmodified reference call the CUDA hot path.

The reference binds the hot-path functions as module globals of their callers
(`from models.modules import *` in first_layer.py:4, second_layer.py:6, third_layer.py:5;
`import tensor_resize` in utils/utils.py:17), so replacing models.modules alone is not enough:
install() rebinds the names inside each caller module.  models/pats.py then runs unmodified.

    import pats_b200.install as inst
    inst.install()            # after the reference's `models` / `utils` packages are importable
    ...
    inst.uninstall()
"""
from __future__ import annotations

import importlib
import sys

from . import layers as _layers
from . import modules as _modules
from . import tensor_resize as _tensor_resize
from . import utils as _utils

# (reference module, attribute) -> replacement
_OT = {
    "log_sinkhorn_iterations": _modules.log_sinkhorn_iterations,
    "log_optimal_transport": _modules.log_optimal_transport,
    "log_optimal_transport2": _modules.log_optimal_transport2,
}
_TABLE = {
    "models.modules": dict(_OT),
    "models.first_layer": {**_OT, "Compute_imgs": _utils.Compute_imgs, "Iterative_expand_matrix": _utils.Iterative_expand_matrix,
                           "split_patches": _utils.split_patches},
    "models.second_layer": {**_OT, "Iterative_expand_matrix": _utils.Iterative_expand_matrix},
    "models.third_layer": dict(_OT),
    "models.pats": {"get_result": _utils.get_result},
    "utils.utils": {"tensor_resize": _tensor_resize, "origin_extract": _utils.origin_extract, "Compute_imgs": _utils.Compute_imgs,
                    "Iterative_expand_matrix": _utils.Iterative_expand_matrix, "get_result": _utils.get_result},
}
# (reference module, class, method) -> replacement: bound methods the layers call through `self`
_METHODS = {
    ("models.second_layer", "SecondLayer", "merge_patches_new"): _layers.merge_patches_new,
    ("models.second_layer", "SecondLayer", "merge_patches_old"): _layers.merge_patches_old,
    ("models.third_layer", "ThirdLayer", "Compute_result"): _layers.Compute_result,
    ("models.first_layer", "FirstLayer", "est_position"): _layers.first_layer_est_position,
    ("models.second_layer", "SecondLayer", "est_position"): _layers.second_layer_est_position,
}
_saved: dict = {}


def install(only=None, fused: bool = False, attention: bool = False) -> list:
    """Rebind the hot-path names; returns the list of (module, name) pairs that were replaced.

    fused=True additionally rebinds `SecondLayer.forward` and `ThirdLayer.forward` to pats_b200.forward's mirrors, which
    reach the pieces of the path that are inline statements in the reference (12 x 12 grid sampling, 8 x 8 window unfold)
    and the composite Sinkhorn -> consumer calls.

    attention=True additionally rebinds `AttentionalGNN.forward` (models/modules.py:126-134, the network in front of every
    matching level: SURVEY.md 8f N3) to pats_b200.gnn.attentional_gnn_forward.  Its 1x1 convolutions then run with FP32-class
    accuracy on the tensor cores, whereas the stock reference on a GPU runs them as single-pass TF32 through cuDNN: the match
    lists agree with the stock GPU run only as far as the stock run agrees with its own CPU execution (tests/test_gpu_gnn.py)."""
    done = []
    sys.modules.setdefault("tensor_resize", _tensor_resize)
    for modname, table in _TABLE.items():
        try:
            mod = importlib.import_module(modname)
        except Exception:
            continue  # that part of the reference is not importable here; nothing to patch
        for name, repl in table.items():
            if only is not None and name not in only:
                continue
            if hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, repl)
                done.append((modname, name))
    methods = dict(_METHODS)
    if fused:
        from .forward import FORWARDS

        methods.update(FORWARDS)
    if attention:
        from .gnn import attentional_gnn_forward

        methods[("models.modules", "AttentionalGNN", "forward")] = attentional_gnn_forward
    for (modname, clsname, meth), repl in methods.items():
        if only is not None and meth not in only:
            continue
        mod = sys.modules.get(modname)
        cls = getattr(mod, clsname, None) if mod is not None else None
        if cls is not None and hasattr(cls, meth):
            _saved.setdefault((modname, clsname + "." + meth), getattr(cls, meth))
            setattr(cls, meth, repl)
            done.append((modname, clsname + "." + meth))
    return done


def uninstall() -> None:
    for (modname, name), orig in list(_saved.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            if "." in name:
                clsname, meth = name.split(".")
                setattr(getattr(mod, clsname), meth, orig)
            else:
                setattr(mod, name, orig)
        del _saved[(modname, name)]
