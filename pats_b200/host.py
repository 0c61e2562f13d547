"""Host-buffer entry points (numpy arrays / CPU tensors in, numpy out): the end-to-end path of the C ABI.

Each call copies its inputs to the device, runs the CUDA kernels and copies the result back
(`*_host` functions of include/pats_b200.h).  Still CUDA-only: without a device the call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _np(a, dtype):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


def log_optimal_transport(scores, alpha, ns, iters: int, out=None):
    scores, ns = _np(scores, np.float32), _np(ns, np.float32)
    b, m, n = scores.shape
    if out is None:
        out = np.empty((b, m + 1, n + 1), np.float32)
    rc = _lib.load().pats_log_optimal_transport_f32_host(scores.ctypes.data, C.c_float(float(alpha)), ns.ctypes.data, b, m, n, int(iters),
                                                         out.ctypes.data)
    _lib.check(rc, "log_optimal_transport (host)")
    return out


def log_optimal_transport2(scores, one, ns, iters: int, out=None):
    scores, ns = _np(scores, np.float32), _np(ns, np.float32)
    b, m, n = scores.shape
    if out is None:
        out = np.empty((b, m, n), np.float32)
    rc = _lib.load().pats_log_optimal_transport2_f32_host(scores.ctypes.data, C.c_float(float(one)), ns.ctypes.data, b, m, n, int(iters),
                                                          out.ctypes.data)
    _lib.check(rc, "log_optimal_transport2 (host)")
    return out


def tensor_resize(input_tensor, bound, out_hw=(96, 96), out=None):
    inp, bound = _np(input_tensor, np.float32), _np(bound, np.int64)
    B, Cc, Hp, Wp = inp.shape
    K = bound.shape[0]
    if out is None:
        out = np.empty((K, Cc, out_hw[0], out_hw[1]), np.float32)
    rc = _lib.load().pats_tensor_resize_f32_host(inp.ctypes.data, B, Cc, Hp, Wp, bound.ctypes.data, K, out_hw[0], out_hw[1], out.ctypes.data)
    _lib.check(rc, "tensor_resize (host)")
    return out
