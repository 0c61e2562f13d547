"""ctypes binding of libpats_b200.so (the C ABI declared in include/pats_b200.h).

PyTorch is used by the callers for device memory and streams only; no torch type crosses
this boundary.  There is NO fallback: if the library is missing or a call fails, a
RuntimeError is raised (the reference raises RuntimeError from its C++ extension too).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("PATS_B200_LIB") or os.path.join(_HERE, "libpats_b200.so")  # override: A/B of two builds of the library
_lib = None

_P = C.c_void_p
_I = C.c_int
_F = C.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPE
SIGNATURES = {
    "pats_version": [],
    "pats_last_error": [],
    "pats_sm_count": [],
    "pats_log_sinkhorn_iterations_f32": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "pats_log_optimal_transport_f32": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "pats_log_optimal_transport2_f32": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "pats_sinkhorn_kernel_kind": [_I, _I],
    "pats_sinkhorn_grid_ctas_per_problem": [_I],
    "pats_sinkhorn_grid_variant": [_I],
    "pats_sinkhorn_force_generic": [_I],
    "pats_plan_handover": [_I],
    "pats_launch_chaining": [_I],
    "pats_sinkhorn_disable_w65": [_I],
    "pats_sinkhorn_disable_c145": [_I],
    "pats_sinkhorn_cluster_variant": [_I],
    "pats_sinkhorn_fallback_count": [_I],
    "pats_sinkhorn_fixed_point_exit": [_I],
    "pats_sinkhorn_bulk_staging": [_I],
    "pats_sinkhorn_iterations_skipped": [_I],
    "pats_log_optimal_transport_f32_host": [_P, _F, _P, _I, _I, _I, _I, _P],
    "pats_log_optimal_transport2_f32_host": [_P, _F, _P, _I, _I, _I, _I, _P],
    "pats_tensor_resize_f32": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P],
    "pats_tensor_resize_f32_variant": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _I, _P],
    "pats_tensor_resize_f32_host": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "pats_origin_extract": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "pats_compute_bounds_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "pats_compute_imgs": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "pats_iterative_expand_matrix_f32": [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "pats_est_nomatching_f32": [_P, _I, _I, _I, _I, _P, _P, _P],
    "pats_merge_patches": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "pats_argsort9_first_f64": [_P, _I, _I, _P, _P],
    "pats_get_result_f32": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P, C.c_longlong, _P, _P, _P],
    "pats_third_compute_result_f32": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    "pats_third_result_from_log_f32": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    "pats_correlation_f32": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "pats_gnn_raw_floats": [_I, _I],
    "pats_gnn_packed_floats": [_I, _I],
    "pats_gnn_workspace_floats": [_I, _I, _I],
    "pats_gnn_pack_f32": [_P, _I, _I, _I, _F, _P, _P],
    "pats_attentional_gnn_f32": [_P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, C.c_longlong, _P],
    "pats_gnn_pack_train_f32": [_P, _I, _I, _I, _P, _P],
    "pats_attentional_gnn_train_f32": [_P, _P, _I, _I, _I, _P, _P, _P, _F, _F, _P, _I, _I, _P, _P, _P, C.c_longlong, _P],
    "pats_gnn_precision": [_I],
    "pats_gnn_attention_variant": [_I],
    "pats_gnn_gemm_variant": [_I],
    "pats_grid_sample12_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "pats_third_unfold_f32": [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P],
    "pats_est_position_f32": [_P, _P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pats_second_layer_match_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pats_third_layer_match_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
}
_RESTYPE = {"pats_last_error": C.c_char_p, "pats_gnn_raw_floats": C.c_longlong, "pats_gnn_packed_floats": C.c_longlong, "pats_gnn_workspace_floats": C.c_longlong, "pats_gnn_precision": None, "pats_gnn_attention_variant": None, "pats_gnn_gemm_variant": None, "pats_sinkhorn_bulk_staging": None, "pats_sinkhorn_fixed_point_exit": None, "pats_sinkhorn_iterations_skipped": C.c_longlong, "pats_sinkhorn_force_generic": None, "pats_plan_handover": None, "pats_launch_chaining": None, "pats_sinkhorn_disable_w65": None, "pats_sinkhorn_disable_c145": None, "pats_sinkhorn_cluster_variant": None, "pats_sinkhorn_grid_ctas_per_problem": None, "pats_sinkhorn_grid_variant": None}


def library_path() -> str:
    return _SO


def load():
    """Load the CUDA library (built by pats_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(
            f"pats_b200: {_SO} is missing. The hot path is CUDA-only (no CPU fallback); "
            "build it with `python -m pats_b200.build` (needs nvcc)."
        )
    lib = C.CDLL(_SO)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pats_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"pats_b200.{what} failed (code {rc}): {msg}")
