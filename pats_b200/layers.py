"""Layer-level pieces of the hot path with the reference's method names.

    merge_patches_new / merge_patches_old      models/second_layer.py:189-238 / :137-186   (SecondLayer methods)
    Compute_result                             models/third_layer.py:184-217               (ThirdLayer method)
They take `self` first so they can be bound over the reference's methods (pats_b200.install).
"""
from __future__ import annotations

import torch

from . import _lib
from ._torchutil import cuda_f32, stream_ptr

__all__ = ["merge_patches_new", "merge_patches_old", "Compute_result", "third_compute_result"]


# How ties of the reference's unstable `torch.argsort(...)[..., 0]` (second_layer.py:169 / :230) are resolved: "cuda" = as ATen's
# CUDA kernel does (default: the reference runs on CUDA tensors), "first" = first minimum, as ATen's CPU kernel does (what
# fixtures recorded from a CPU run of the reference hold).  See kArgsortTie9 in csrc/regroup.cu.
MERGE_TIE_BREAK = "cuda"


def _merge(merge_new, patch_num, trust_score, original_image_shape, if_nomatching1_L1, if_nomatching1_L2, scores_back):
    dev = trust_score.device
    if not trust_score.is_cuda:
        raise RuntimeError("merge_patches: tensors are on the CPU; pats_b200 is CUDA-only (no CPU fallback)")
    height, width = int(original_image_shape[0]) // 32, int(original_image_shape[1]) // 32
    B = if_nomatching1_L1.shape[0]
    P = int(if_nomatching1_L2.shape[0])
    # the reference mutates trust_score / if_nomatching1_L2 / scores_back in place: work on the caller's storage
    t = trust_score if (trust_score.is_contiguous() and trust_score.dtype == torch.float32) else trust_score.float().contiguous()
    f = if_nomatching1_L2 if (if_nomatching1_L2.is_contiguous() and if_nomatching1_L2.dtype == torch.bool) else if_nomatching1_L2.bool().contiguous()
    sb = scores_back if (scores_back.is_contiguous() and scores_back.dtype == torch.float64) else scores_back.double().contiguous()
    nm1 = if_nomatching1_L1.to(torch.uint8).contiguous()
    out = torch.empty((P, 144), dtype=torch.bool, device=dev)
    ws = torch.empty(2 * B * height * width + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_merge_patches((1 if merge_new else 0) | (2 if MERGE_TIE_BREAK == "first" else 0), t.data_ptr(), nm1.data_ptr(), f.data_ptr(), sb.data_ptr(), B, height, width, P,
                                            out.data_ptr(), ws.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "merge_patches")
    if t is not trust_score:
        trust_score.copy_(t)
    if f is not if_nomatching1_L2:
        if_nomatching1_L2.copy_(f)
    if sb is not scores_back:
        scores_back.copy_(sb)
    return out, sb


def merge_patches_new(self, patch_num, trust_score, original_image_shape, if_nomatching1_L1, if_nomatching1_L2, scores_back):
    """models/second_layer.py:189-238; scores_back is carried across chunks."""
    return _merge(True, patch_num, trust_score, original_image_shape, if_nomatching1_L1, if_nomatching1_L2, scores_back)


def merge_patches_old(self, patch_num, trust_score, original_image_shape, if_nomatching1_L1, if_nomatching1_L2, scores_back):
    """models/second_layer.py:137-186; returns zeroed scores_back."""
    return _merge(False, patch_num, trust_score, original_image_shape, if_nomatching1_L1, if_nomatching1_L2, scores_back)


def third_compute_result(scores, scale_x, scale_y, p_s, p_t):
    """(mkpts0_f, mkpts1_f, if_matching1) for K level-3 problems: third_layer.py:184-217 and the label test :166-167."""
    scores = cuda_f32(scores, "scores")
    K = scores.shape[0]
    if scores.shape[1:] != (65, 65):
        raise ValueError(f"third layer plans are [K,65,65] (W=8), got {tuple(scores.shape)}")
    dev = scores.device
    sx = cuda_f32(scale_x, "scale_x").reshape(K, 64)
    sy = cuda_f32(scale_y, "scale_y").reshape(K, 64)
    ps = p_s.to(device=dev, dtype=torch.int64).contiguous()
    pt = p_t.to(device=dev, dtype=torch.int64).contiguous()
    m0 = torch.empty((K, 16, 2), dtype=torch.float32, device=dev)
    m1 = torch.empty_like(m0)
    im = torch.empty((K, 16), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_third_compute_result_f32(scores.data_ptr(), sx.data_ptr(), sy.data_ptr(), ps.data_ptr(), pt.data_ptr(), K, m0.data_ptr(),
                                                       m1.data_ptr(), im.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "Compute_result")
    return m0, m1, im


def Compute_result(self, scores, W, T, scale_x, scale_y, p_s, p_t, device):
    """ThirdLayer.Compute_result (third_layer.py:184-217).  whole_loss, the third return value, is discarded by the only
    caller (third_layer.py:160) and is returned as zeros."""
    if W != 8 or T != 5:
        raise NotImplementedError("Compute_result: the reference fixes W=8, T=5 (third_layer.py:107-110)")
    m0, m1, _ = third_compute_result(scores, scale_x, scale_y, p_s, p_t)
    return m0, m1, torch.zeros((scores.shape[0], 16), dtype=torch.float32, device=scores.device)


def est_position(scores, scale_x, scale_y, grid_h: int, grid_w: int, iter_num: int, lower_bound: float, *, return_extra=False):
    """Fused est_position (first_layer.py:159-178 with scale_x = scale_y = scale_src, iter_num=15, lower_bound=1e-5;
    second_layer.py:240-259 with iter_num=8, lower_bound=1e-3) from the LOG-domain plan `scores` [b,n+1,n+1].

    Returns (trust_score [b,n], average_point [b,n,2], x_scale [b,n], y_scale [b,n], if_nomatching1 [b,n], if_nomatching2 [b,n]).
    """
    scores = cuda_f32(scores, "scores")
    b, M, N = scores.shape
    n = grid_h * grid_w
    if M != n + 1 or N != n + 1:
        raise ValueError(f"est_position: plan {tuple(scores.shape)} is not [b,{n + 1},{n + 1}] for a {grid_h}x{grid_w} grid")
    dev = scores.device
    sx = cuda_f32(scale_x, "scale_x").reshape(b, n)
    sy = cuda_f32(scale_y, "scale_y").reshape(b, n)
    trust = torch.empty((b, n), dtype=torch.float32, device=dev)
    avg = torch.empty((b, n, 2), dtype=torch.float32, device=dev)
    xs, ys, core = torch.empty_like(trust), torch.empty_like(trust), torch.empty_like(trust)
    nm1 = torch.empty((b, n), dtype=torch.bool, device=dev)
    nm2 = torch.empty_like(nm1)
    bound = torch.empty((b, n, 4), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_est_position_f32(scores.data_ptr(), sx.data_ptr(), sy.data_ptr(), b, grid_h, grid_w, float(lower_bound), int(iter_num),
                                               trust.data_ptr(), avg.data_ptr(), xs.data_ptr(), ys.data_ptr(), nm1.data_ptr(), nm2.data_ptr(),
                                               core.data_ptr(), bound.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "est_position")
    out = (trust, avg, xs, ys, nm1, nm2)
    return out + (core, bound) if return_extra else out


def first_layer_est_position(self, scores, scale_src, image_shape, patch_scale):
    """FirstLayer.est_position (first_layer.py:159-178)."""
    H, W = image_shape
    return est_position(scores, scale_src, scale_src, H // patch_scale, W // patch_scale, 15, 1e-5)


def second_layer_est_position(self, scores, scale_x, scale_y, image_shape, patch_scale):
    """SecondLayer.est_position (second_layer.py:240-259)."""
    H, W = image_shape
    return est_position(scores, scale_x, scale_y, H // patch_scale, W // patch_scale, 8, 1e-3)


def third_result_from_log(Z, scale_x, scale_y, p_s, p_t):
    """third_compute_result on the log-domain plan (exp fused into the load; third_layer.py:158-160)."""
    Z = cuda_f32(Z, "Z")
    K = Z.shape[0]
    dev = Z.device
    sx = cuda_f32(scale_x, "scale_x").reshape(K, 64)
    sy = cuda_f32(scale_y, "scale_y").reshape(K, 64)
    ps = p_s.to(device=dev, dtype=torch.int64).contiguous()
    pt = p_t.to(device=dev, dtype=torch.int64).contiguous()
    m0 = torch.empty((K, 16, 2), dtype=torch.float32, device=dev)
    m1 = torch.empty_like(m0)
    im = torch.empty((K, 16), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_third_result_from_log_f32(Z.data_ptr(), sx.data_ptr(), sy.data_ptr(), ps.data_ptr(), pt.data_ptr(), K, m0.data_ptr(),
                                                        m1.data_ptr(), im.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "third_result_from_log")
    return m0, m1, im


_EDGE_CACHE: dict = {}


def second_layer_match(scores, one, scale, scale_x, scale_y, iters: int = 100, outdoor: bool = True, grid: int = 12, *, return_extra=False):
    """The matching block of SecondLayer.forward (second_layer.py:103-116) in one call:

        scores = log_optimal_transport2(scores, one, scale, iters)
        scores[:, :, -1] += log(2 or 3); scores[:, -1, :] += log(2 or 3)          # outdoor / indoor, :108-112
        trust_score, pts, x_scale, y_scale, if_nomatching1, if_nomatching2 = est_position(scores, scale_x, scale_y, ...)

    `scores` is the already scaled input (0.1 * einsum / sqrt(d)), [P, grid^2+1, grid^2+1].  Returns (scores_out, trust_score,
    pts, x_scale, y_scale, if_nomatching1, if_nomatching2); scores_out is what the reference keeps as 'scores'.  The area
    expansion starts on finished problems while the Sinkhorn kernel's last wave is still running."""
    import math

    scores = cuda_f32(scores, "scores")
    b, M, N = scores.shape
    n = grid * grid
    if M != n + 1 or N != n + 1:
        raise ValueError(f"second_layer_match: scores {tuple(scores.shape)} is not [b,{n + 1},{n + 1}]")
    dev = scores.device
    ns = cuda_f32(scale, "scale")
    if ns.numel() != b * n:
        raise ValueError(f"scale must hold b*{n} values, got {tuple(ns.shape)}")
    sx = cuda_f32(scale_x, "scale_x").reshape(b, n)
    sy = cuda_f32(scale_y, "scale_y").reshape(b, n)
    from ._torchutil import scalar_on

    one_t = scalar_on(dev, one, "one")
    # f32(log(f32(one) * k)) as torch.log(self.one * 2) computes it.  A number costs nothing; a CUDA tensor (the reference's
    # nn.Parameter `one`) is read back once per distinct (storage, version) and cached -- no device sync in steady state.
    k = 2.0 if outdoor else 3.0
    if isinstance(one, torch.Tensor):
        key = (one.data_ptr(), one._version, k)
        edge = _EDGE_CACHE.get(key)
        if edge is None:
            edge = float(torch.log(one.detach().float().cpu() * k))
            _EDGE_CACHE.clear()
            _EDGE_CACHE[key] = edge
    else:
        edge = float(torch.log(torch.tensor(float(one), dtype=torch.float32) * k))
    Z = torch.empty_like(scores)
    trust = torch.empty((b, n), dtype=torch.float32, device=dev)
    avg = torch.empty((b, n, 2), dtype=torch.float32, device=dev)
    xs, ys, core = torch.empty_like(trust), torch.empty_like(trust), torch.empty_like(trust)
    nm1 = torch.empty((b, n), dtype=torch.bool, device=dev)
    nm2 = torch.empty_like(nm1)
    bound = torch.empty((b, n, 4), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_second_layer_match_f32(scores.data_ptr(), one_t.data_ptr(), ns.data_ptr(), sx.data_ptr(), sy.data_ptr(), b, grid, grid,
                                                     int(iters), edge, 1e-3, 8, Z.data_ptr(), trust.data_ptr(), avg.data_ptr(), xs.data_ptr(),
                                                     ys.data_ptr(), nm1.data_ptr(), nm2.data_ptr(), core.data_ptr(), bound.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "second_layer_match")
    out = (Z, trust, avg, xs, ys, nm1, nm2)
    return out + (core, bound) if return_extra else out


def third_layer_match(scores, one, scale, scale_x, scale_y, p_s, p_t, iters: int = 100):
    """The matching block of ThirdLayer.forward (third_layer.py:158-167) in one call: log_optimal_transport2 on the [K,65,65]
    scores (already 0.1 * einsum / sqrt(128)), exp, Compute_result and the label test.
    Returns (scores_origin [K,65,65], mkpts0_f, mkpts1_f [K,16,2], if_matching1 [K,16])."""
    scores = cuda_f32(scores, "scores")
    K = scores.shape[0]
    if scores.shape[1:] != (65, 65):
        raise ValueError(f"third layer scores are [K,65,65] (W=8), got {tuple(scores.shape)}")
    dev = scores.device
    ns = cuda_f32(scale, "scale")
    if ns.numel() != K * 64:
        raise ValueError(f"scale must hold K*64 values, got {tuple(ns.shape)}")
    sx = cuda_f32(scale_x, "scale_x").reshape(K, 64)
    sy = cuda_f32(scale_y, "scale_y").reshape(K, 64)
    ps = p_s.to(device=dev, dtype=torch.int64).contiguous()
    pt = p_t.to(device=dev, dtype=torch.int64).contiguous()
    from ._torchutil import scalar_on

    one_t = scalar_on(dev, one, "one")
    Z = torch.empty_like(scores)
    m0 = torch.empty((K, 16, 2), dtype=torch.float32, device=dev)
    m1 = torch.empty_like(m0)
    im = torch.empty((K, 16), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_third_layer_match_f32(scores.data_ptr(), one_t.data_ptr(), ns.data_ptr(), sx.data_ptr(), sy.data_ptr(), ps.data_ptr(),
                                                    pt.data_ptr(), K, int(iters), Z.data_ptr(), m0.data_ptr(), m1.data_ptr(), im.data_ptr(),
                                                    stream_ptr(dev))
    _lib.check(rc, "third_layer_match")
    return Z, m0, m1, im


__all__ += ["est_position", "first_layer_est_position", "second_layer_est_position", "third_result_from_log", "second_layer_match",
            "third_layer_match"]


def grid_sample12(maps, row_num: int = 12):
    """second_layer.py:71-80: sample the three stem maps [N,64,48,48], [N,64,24,24], [N,128,12,12] on the 12x12 grid
    (levels 0/1 through AvgPool2d(2, stride=1, padding=1)) and concatenate along channels -> [N,256,144]."""
    f0, f1, f2 = (cuda_f32(m, f"maps[{i}]") for i, m in enumerate(maps))
    N = f0.shape[0]
    R = row_num
    if f0.shape[2:] != (4 * R, 4 * R) or f1.shape[2:] != (2 * R, 2 * R) or f2.shape[2:] != (R, R):
        raise ValueError("grid_sample12: maps must be [N,C0,4R,4R], [N,C1,2R,2R], [N,C2,R,R]")
    out = torch.empty((N, f0.shape[1] + f1.shape[1] + f2.shape[1], R * R), dtype=torch.float32, device=f0.device)
    with torch.cuda.device(f0.device):
        rc = _lib.load().pats_grid_sample12_f32(f0.data_ptr(), f1.data_ptr(), f2.data_ptr(), N, f0.shape[1], f1.shape[1], f2.shape[1], R,
                                                out.data_ptr(), stream_ptr(f0.device))
    _lib.check(rc, "grid_sample12")
    return out


def third_unfold(feat, mkpts_c, b_ids, kenc_out, rubbish, mkpts0_c, clamp96: bool, *, check: bool = True):
    """third_layer.py:119-146 for one image side: 8x8 window features around every point + kenc(kpts) + rubbish token
    -> [K,C,65] (the GNN input).  mkpts_c / mkpts0_c are the raw (x,y) points of third_layer.forward's arguments."""
    feat = cuda_f32(feat, "feat")
    P, Cc, M, M2 = feat.shape
    dev = feat.device
    mk = cuda_f32(mkpts_c, "mkpts_c").reshape(-1, 2)
    K = mk.shape[0]
    mk0 = cuda_f32(mkpts0_c, "mkpts0_c").reshape(K, 2)
    b = cuda_f32(b_ids, "b_ids").reshape(K)
    kenc = cuda_f32(kenc_out, "kenc_out").reshape(Cc, 64)
    rub = cuda_f32(rubbish, "rubbish").reshape(P, Cc, 144)
    out = torch.empty((K, Cc, 65), dtype=torch.float32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_third_unfold_f32(feat.data_ptr(), P, Cc, M, mk.data_ptr(), b.data_ptr(), K, 1 if clamp96 else 0, kenc.data_ptr(),
                                               rub.data_ptr(), mk0.data_ptr(), out.data_ptr(), bad.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "third_unfold")
    if check and K > 0 and int(bad.item()):
        raise IndexError("third_unfold: a window's gather index lies outside the feature tensor")
    return out


__all__ += ["grid_sample12", "third_unfold"]


def correlation(desc0, desc1, scale: float):
    """scale * einsum('bdn,bdm->bnm', desc0, desc1): the descriptor correlation in front of every Sinkhorn solve
    (first_layer.py:110-114, second_layer.py:100-104, third_layer.py:156-158 with scale = 0.1 / sqrt(d)), on the tcgen05 tensor cores
    with FP32-level accuracy (3xTF32).  desc0 [b,d,n], desc1 [b,d,m] -> [b,n,m]."""
    d0 = cuda_f32(desc0, "desc0")
    d1 = cuda_f32(desc1, "desc1")
    if d0.dim() != 3 or d1.dim() != 3 or d0.shape[:2] != d1.shape[:2]:
        raise ValueError(f"correlation: expected [b,d,n] and [b,d,m], got {tuple(d0.shape)} and {tuple(d1.shape)}")
    b, d, n = d0.shape
    m = d1.shape[2]
    out = torch.empty((b, n, m), dtype=torch.float32, device=d0.device)
    with torch.cuda.device(d0.device):
        rc = _lib.load().pats_correlation_f32(d0.data_ptr(), d1.data_ptr(), b, d, n, m, float(scale), out.data_ptr(), stream_ptr(d0.device))
    _lib.check(rc, "correlation")
    return out


__all__ += ["correlation"]
