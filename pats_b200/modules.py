"""Sinkhorn / optimal-transport ops with the reference's names and signatures.

Mirrors models/modules.py:137-182 of zju3dv/pats:
    log_sinkhorn_iterations(Z, log_mu, log_nu, iters)      :137-143
    log_optimal_transport(scores, alpha, ns, iters)        :145-162
    log_optimal_transport2(scores, one, ns, iters)         :165-182
Inputs are borrowed and never modified; the result is a fresh contiguous f32 CUDA tensor on the
input's device (second_layer.py:108-112 mutates it in place).  Runs on the current CUDA stream,
no host synchronisation, no autograd (callers are under torch.no_grad, evaluate.py:20).
"""
from __future__ import annotations

import torch

from . import _lib
from ._torchutil import cuda_f32, scalar_on, stream_ptr

__all__ = ["log_sinkhorn_iterations", "log_optimal_transport", "log_optimal_transport2"]


def log_sinkhorn_iterations(Z: torch.Tensor, log_mu: torch.Tensor, log_nu: torch.Tensor, iters: int) -> torch.Tensor:
    """Perform Sinkhorn Normalization in Log-space for stability (models/modules.py:137)."""
    Z = cuda_f32(Z, "Z")
    if Z.dim() != 3:
        raise ValueError(f"Z must be [b,M,N], got {tuple(Z.shape)}")
    b, M, N = Z.shape
    log_mu = cuda_f32(log_mu, "log_mu").reshape(b, M)
    log_nu = cuda_f32(log_nu, "log_nu").reshape(b, N)
    out = torch.empty_like(Z)
    with torch.cuda.device(Z.device):
        rc = _lib.load().pats_log_sinkhorn_iterations_f32(Z.data_ptr(), log_mu.data_ptr(), log_nu.data_ptr(), b, M, N, int(iters),
                                                          out.data_ptr(), stream_ptr(Z.device))
    _lib.check(rc, "log_sinkhorn_iterations")
    return out


def log_optimal_transport(scores: torch.Tensor, alpha, ns: torch.Tensor, iters: int) -> torch.Tensor:
    """Differentiable-OT forward in log space with dustbin augmentation (models/modules.py:145).

    scores [b,m,n], alpha 0-dim tensor (or number), ns [b,1,n]  ->  [b,m+1,n+1]
    """
    scores = cuda_f32(scores, "scores")
    if scores.dim() != 3:
        raise ValueError(f"scores must be [b,m,n], got {tuple(scores.shape)}")
    b, m, n = scores.shape
    ns = cuda_f32(ns, "ns")
    if ns.numel() != b * n:
        raise ValueError(f"ns must hold b*n = {b * n} values, got {tuple(ns.shape)}")
    alpha_t = scalar_on(scores.device, alpha, "alpha")
    out = torch.empty((b, m + 1, n + 1), dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().pats_log_optimal_transport_f32(scores.data_ptr(), alpha_t.data_ptr(), ns.data_ptr(), b, m, n, int(iters),
                                                        out.data_ptr(), stream_ptr(scores.device))
    _lib.check(rc, "log_optimal_transport")
    return out


def log_optimal_transport2(scores: torch.Tensor, one, ns: torch.Tensor, iters: int) -> torch.Tensor:
    """OT forward where the dustbin already is the last row / column (models/modules.py:165).

    scores [b,m,n], one 0-dim tensor (or number), ns [b,1,n-1]  ->  [b,m,n]
    """
    scores = cuda_f32(scores, "scores")
    if scores.dim() != 3:
        raise ValueError(f"scores must be [b,m,n], got {tuple(scores.shape)}")
    b, m, n = scores.shape
    ns = cuda_f32(ns, "ns")
    if ns.numel() != b * (n - 1):
        raise ValueError(f"ns must hold b*(n-1) = {b * (n - 1)} values, got {tuple(ns.shape)}")
    one_t = scalar_on(scores.device, one, "one")
    out = torch.empty((b, m, n), dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().pats_log_optimal_transport2_f32(scores.data_ptr(), one_t.data_ptr(), ns.data_ptr(), b, m, n, int(iters),
                                                         out.data_ptr(), stream_ptr(scores.device))
    _lib.check(rc, "log_optimal_transport2")
    return out
