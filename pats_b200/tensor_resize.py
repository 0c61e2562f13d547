"""Drop-in for the reference's pybind11 module `tensor_resize` (setup/library.cpp:92-93).

    tensor_resize.tensor_resize(input_tensor, bound) -> Tensor        "feature resize"

input [B,C,Hp,Wp] f32 CUDA, bound [K,5] int64 rows (y0,y1,x0,x1,img*10000+patch) -> [K,C,96,96] f32.
One kernel launch; the bounds are read on the device (the reference does 5 .item() syncs per
patch, library.cpp:55-59).  Like the reference (c10::Error from narrow/upsample -> RuntimeError)
an empty or out-of-range crop raises RuntimeError; that check costs the call's only host sync and
can be disabled with `tensor_resize.CHECK_BOUNDS = False`.
"""
from __future__ import annotations

import torch

from . import _lib
from ._torchutil import cuda_f32, stream_ptr

CHECK_BOUNDS = True
OUT_HW = (96, 96)  # library.cpp:50-52: patch_shape * 3


def tensor_resize(input_tensor: torch.Tensor, bound: torch.Tensor, *, variant: int | None = None) -> torch.Tensor:
    """feature resize"""
    inp = cuda_f32(input_tensor, "input_tensor")
    if inp.dim() != 4:
        raise ValueError(f"input_tensor must be [B,C,H,W], got {tuple(inp.shape)}")
    if bound.dim() != 2 or bound.shape[1] != 5:
        raise ValueError(f"bound must be [K,5], got {tuple(bound.shape)}")
    bound = bound.to(device=inp.device, dtype=torch.int64).contiguous()
    B, Cc, Hp, Wp = inp.shape
    K = bound.shape[0]
    out = torch.empty((K, Cc, OUT_HW[0], OUT_HW[1]), dtype=torch.float32, device=inp.device)
    bad = torch.zeros(1, dtype=torch.int32, device=inp.device) if CHECK_BOUNDS else None
    bad_ptr = bad.data_ptr() if bad is not None else None
    with torch.cuda.device(inp.device):
        if variant is None:  # shipping lerp recipe (bit-identical to ATen's CUDA bilinear kernel)
            rc = _lib.load().pats_tensor_resize_f32(inp.data_ptr(), B, Cc, Hp, Wp, bound.data_ptr(), K, OUT_HW[0], OUT_HW[1], out.data_ptr(),
                                                    bad_ptr, stream_ptr(inp.device))
        else:  # explicit rounding recipe, tests only
            rc = _lib.load().pats_tensor_resize_f32_variant(inp.data_ptr(), B, Cc, Hp, Wp, bound.data_ptr(), K, OUT_HW[0], OUT_HW[1],
                                                            out.data_ptr(), bad_ptr, int(variant), stream_ptr(inp.device))
    _lib.check(rc, "tensor_resize")
    if bad is not None and K > 0:
        nbad = int(bad.item())
        if nbad:
            raise RuntimeError(f"tensor_resize: {nbad} bound row(s) describe an empty or out-of-range crop")
    return out
