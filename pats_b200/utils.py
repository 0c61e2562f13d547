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
