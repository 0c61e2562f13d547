"""Patch-subdivision helpers with the reference's names and signatures (utils/utils.py of zju3dv/pats).

    origin_extract(left, patch_scale, width, height)                         utils/utils.py:1300-1318
    Compute_imgs(x_scale, y_scale, average_point, if_nomatching, left, right, ...)   :1343-1393
    compute_bounds(...)   the bound / scale arithmetic of Compute_imgs on its own  :1357-1372
Everything runs in libpats_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib
from ._torchutil import cuda_f32, stream_ptr

__all__ = ["origin_extract", "Compute_imgs", "compute_bounds"]


def _require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: pats_b200 is CUDA-only (no CPU fallback)")
    return t.contiguous()


def origin_extract(left: torch.Tensor, patch_scale: int, width: int, height: int, if_swap: bool = False, average_point=None):
    """3x3-patch window around every coarse patch of the one-patch-padded image (utils/utils.py:1300).

    left [B,C,ps*(height+2),ps*(width+2)] (any dtype) -> [B,C,height*width,3ps,3ps]; bit-exact copy.
    """
    if if_swap:
        raise NotImplementedError("origin_extract(if_swap=True) is dead code in the reference (never called)")
    left = _require_cuda(left, "left")
    B, Cc, Hs, Ws = left.shape
    if Hs != patch_scale * (height + 2) or Ws != patch_scale * (width + 2):
        raise ValueError(f"left {tuple(left.shape)} is not the image padded by one {patch_scale}-px patch for a {height}x{width} grid")
    out = torch.empty((B, Cc, height * width, 3 * patch_scale, 3 * patch_scale), dtype=left.dtype, device=left.device)
    with torch.cuda.device(left.device):
        rc = _lib.load().pats_origin_extract(left.data_ptr(), left.element_size(), B, Cc, height, width, patch_scale, out.data_ptr(),
                                             stream_ptr(left.device))
    _lib.check(rc, "origin_extract")
    return out


def compute_bounds(x_scale, y_scale, average_point, height: int, width: int, patch_scale: int = 32, margin: int = 128):
    """Crop bounds / re-derived scales of Compute_imgs (utils/utils.py:1357-1372,1380-1381)."""
    x_scale, y_scale, average_point = cuda_f32(x_scale, "x_scale"), cuda_f32(y_scale, "y_scale"), cuda_f32(average_point, "average_point")
    B, n = x_scale.shape
    dev = x_scale.device
    bound = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
    xs = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
    ys = torch.empty_like(xs)
    avg = torch.empty_like(xs)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_compute_bounds_f32(x_scale.data_ptr(), y_scale.data_ptr(), average_point.data_ptr(), B, height, width,
                                                 patch_scale, margin, bound.data_ptr(), xs.data_ptr(), ys.data_ptr(), avg.data_ptr(),
                                                 stream_ptr(dev))
    _lib.check(rc, "compute_bounds")
    return bound, xs, ys, avg


def Compute_imgs(x_scale, y_scale, average_point, if_nomatching, left, right, sequence_num=0, output_path=None, if_view=False,
                 margin=128, width=20, height=15, patch_scale=32, *, return_bound=False):
    """Subdivide: left 96x96 windows + right crop/resized patches of the matched coarse patches (utils/utils.py:1343).

    left,right [B,H,W,3] uint8 (as evaluate.py:26-27 feeds them) or float.  Returns
    (new_left [P,96,96,3] dtype of left, new_right [P,96,96,3] f32, x_scale_new [B,n,2], y_scale_new [B,n,2],
    average_new [B,n,2]).  One fused pass: no F.pad, no index temporaries, no per-patch host syncs; the only
    synchronisation is reading P (the reference synchronises for its boolean-mask indexing as well).
    """
    if if_view:
        raise NotImplementedError("if_view (cv2 debug dumps) is not part of the hot path")
    x_scale, y_scale, average_point = cuda_f32(x_scale, "x_scale"), cuda_f32(y_scale, "y_scale"), cuda_f32(average_point, "average_point")
    dev = x_scale.device
    left, right = _require_cuda(left, "left"), _require_cuda(right, "right")
    if left.dtype != right.dtype:
        right = right.to(left.dtype)
    if left.dtype not in (torch.uint8, torch.float32):
        left, right = left.float(), right.float()
    B, H, W, ch = left.shape
    if ch != 3 or H != patch_scale * height or W != patch_scale * width or right.shape != left.shape:
        raise ValueError(f"images {tuple(left.shape)}/{tuple(right.shape)} do not match a {height}x{width} grid of {patch_scale}-px patches")
    n = height * width
    nm = if_nomatching.to(device=dev).reshape(B, n).to(torch.uint8).contiguous()
    P = int((nm == 0).sum().item())
    ps3 = 3 * patch_scale
    new_left = torch.empty((P, ps3, ps3, 3), dtype=left.dtype, device=dev)
    new_right = torch.empty((P, 3, ps3, ps3), dtype=torch.float32, device=dev)
    bound5 = torch.empty((P, 5), dtype=torch.int64, device=dev)
    xs = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
    ys = torch.empty_like(xs)
    avg = torch.empty_like(xs)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)  # [count, bad_rows]
    with torch.cuda.device(dev):
        rc = _lib.load().pats_compute_imgs(x_scale.data_ptr(), y_scale.data_ptr(), average_point.data_ptr(), nm.data_ptr(), left.data_ptr(),
                                           right.data_ptr(), left.element_size(), B, height, width, patch_scale, margin, new_left.data_ptr(),
                                           new_right.data_ptr(), bound5.data_ptr(), xs.data_ptr(), ys.data_ptr(), avg.data_ptr(), P,
                                           meta.data_ptr(), meta.data_ptr() + 4, stream_ptr(dev))
    _lib.check(rc, "Compute_imgs")
    from . import tensor_resize as _tr

    if _tr.CHECK_BOUNDS and P > 0:
        nbad = int(meta[1].item())
        if nbad:
            raise RuntimeError(f"Compute_imgs: {nbad} matched patch(es) have an empty or out-of-range crop")
    out = (new_left, new_right.permute(0, 2, 3, 1), xs, ys, avg)
    return out + (bound5,) if return_bound else out


# ---------------------------------------------------------------------------------------------------------
# area expansion / match assembly
# ---------------------------------------------------------------------------------------------------------
def _grid_dims(limitation, ranges, positions, width, height):
    """The reference rebuilds the grid from `ranges` / `positions` (utils.py:1181) and reads limitation[3] on the
    device; both callers also pass the true grid as keywords (first_layer.py:175-176, second_layer.py:255-257),
    which avoids a host sync.  Fall back to reading `limitation` when the keywords do not describe the tensors."""
    n = positions.shape[0]
    if width * height == n and ranges.shape[0] == max(width, height):
        return int(height), int(width)
    lim = [int(v) for v in limitation.tolist()]
    return lim[1], lim[3]


def Iterative_expand_matrix(scores_in, scalex, scaley, limitation, ranges, positions, lower_bound=1e-3, upper_bound=1e7, iter_num=15,
                            width=20, height=15, type="distance", *, return_nomatching=False):
    """Grow the matched area box around every source patch's best target cell (utils/utils.py:1179-1297).

    scores_in [b,m+1,n+1] = exp(Z); scalex, scaley [b,n,1].  Returns (whole_cost [b,m], core_cost [b,m],
    average_point [b,m,2], x_scale [b,m], y_scale [b,m], bound [b,m,4] int64).
    """
    scores_in = cuda_f32(scores_in, "scores_in")
    b, M, N = scores_in.shape
    m, n = M - 1, N - 1
    grid_h, grid_w = _grid_dims(limitation, ranges, positions, width, height)
    if grid_h * grid_w != n:
        raise ValueError(f"grid {grid_h}x{grid_w} does not match {n} target cells")
    sx = cuda_f32(scalex, "scalex").reshape(b, n)
    sy = cuda_f32(scaley, "scaley").reshape(b, n)
    dev = scores_in.device
    whole = torch.empty((b, m), dtype=torch.float32, device=dev)
    core = torch.empty_like(whole)
    avg = torch.empty((b, m, 2), dtype=torch.float32, device=dev)
    xs = torch.empty_like(whole)
    ys = torch.empty_like(whole)
    bound = torch.empty((b, m, 4), dtype=torch.int64, device=dev)
    nm = torch.empty((b, m), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_iterative_expand_matrix_f32(scores_in.data_ptr(), sx.data_ptr(), sy.data_ptr(), b, m, grid_h, grid_w,
                                                          float(lower_bound), int(iter_num), whole.data_ptr(), core.data_ptr(), avg.data_ptr(),
                                                          xs.data_ptr(), ys.data_ptr(), bound.data_ptr(), nm.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "Iterative_expand_matrix")
    out = (whole, core, avg, xs, ys, bound)
    return out + (nm,) if return_nomatching else out


def est_nomatching(scores, dust: int):
    """The two masks of est_position (first_layer.py:162-167): (argmax over columns == dust)[:, :-1], (argmax over rows == dust)[:, :-1]."""
    scores = cuda_f32(scores, "scores")
    b, M, N = scores.shape
    nm1 = torch.empty((b, M - 1), dtype=torch.bool, device=scores.device)
    nm2 = torch.empty((b, N - 1), dtype=torch.bool, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().pats_est_nomatching_f32(scores.data_ptr(), b, M, N, int(dust), nm1.data_ptr(), nm2.data_ptr(), stream_ptr(scores.device))
    _lib.check(rc, "est_nomatching")
    return nm1, nm2


def get_result(batch_size, if_nomatching, average_point, scale, patch_size, left_choice, layer_num=2):
    """Compose level-0 patch geometry with level-1 cell positions into absolute matches (utils/utils.py:189-213).

    Two levels with left_choice all True, as models/pats.py:72-78 calls it.  Returns (matches_l, matches_r) [Kf,2] f32 in the
    row-major order of the reference's boolean-mask indexing.
    """
    if layer_num != 2 or len(if_nomatching) != 2:
        raise NotImplementedError("get_result: the hot path uses exactly two levels (models/pats.py:72-78)")
    nm0 = _require_cuda(if_nomatching[0], "if_nomatching[0]")
    dev = nm0.device
    nm0 = nm0.to(torch.uint8)
    nm1 = _require_cuda(if_nomatching[1], "if_nomatching[1]").to(torch.uint8)
    pt0, pt1 = cuda_f32(average_point[0], "average_point[0]"), cuda_f32(average_point[1], "average_point[1]")
    sc0, sc1 = cuda_f32(scale[0], "scale[0]"), cuda_f32(scale[1], "scale[1]")
    (ps0, h0, w0), (ps1, h1, w1) = [[int(v) for v in s] for s in patch_size]
    B, n0 = nm0.shape
    P, n1 = nm1.shape
    if n0 != h0 * w0 or n1 != h1 * w1:
        raise ValueError("get_result: mask shapes do not match patch_size")
    cap = P * n1
    ml = torch.empty((max(cap, 1), 2), dtype=torch.float32, device=dev)
    mr = torch.empty_like(ml)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(8 * (P + 1) + 4 * (2 * B * n0 + 1 + P) + 8, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pats_get_result_f32(nm0.data_ptr(), pt0.data_ptr(), sc0.data_ptr(), B, ps0, h0, w0, nm1.data_ptr(), pt1.data_ptr(),
                                             sc1.data_ptr(), P, ps1, h1, w1, ml.data_ptr(), mr.data_ptr(), cap, total.data_ptr(),
                                             ws.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "get_result")
    kf = int(total.item())  # the reference synchronises here too (boolean-mask indexing)
    return ml[:kf], mr[:kf]


__all__ += ["Iterative_expand_matrix", "est_nomatching", "get_result"]


def split_patches(sum_cycle, height, width, max_once_used=350):
    """Row-aligned chunking of the matched coarse patches (utils/utils.py:152-181; called at first_layer.py:135).

    Host integer logic in the reference too -- a Python loop that compares 0-dim CUDA tensors (one device sync per
    row).  Here `sum_cycle` (inclusive cumsum of the matched mask) is copied to the host ONCE and the same decisions
    are taken on plain ints; the returned sets hold ints, which every consumer accepts (tensor comparisons at
    first_layer.py:138-139, `!= 0` and negative slicing at models/pats.py:38-39).  Python's negative indexing of the
    reference (`sum_cycle[i*width-1]` with i == 0 reads the last element) is kept.
    """
    sc = [int(v) for v in sum_cycle.detach().reshape(-1).tolist()]
    cycle_num, second_layer_set, third_layer_set = 0, [], []
    last_second_line = last_third_line = 0
    for i in range(height):
        num = sc[(i + 1) * width - 1]
        if num > max_once_used * (cycle_num + 1):
            origin_num = 0 if last_second_line == 0 else sc[last_second_line * width - 1]
            cycle_num += 1
            second_layer_set.append([origin_num, num])
            third_layer_set.append([sc[last_third_line * width] - origin_num, num - sc[i * width - 1]])
            last_second_line, last_third_line = i, i + 1
    origin_num = 0 if last_second_line == 0 else sc[last_second_line * width - 1]
    cycle_num += 1
    second_layer_set.append([origin_num, height * width])
    end_num = origin_num if last_third_line == height else sc[last_third_line * width]
    third_layer_set.append([end_num - origin_num, 0])
    return cycle_num, second_layer_set, third_layer_set


__all__ += ["split_patches"]
