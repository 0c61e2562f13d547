"""Small helpers shared by the torch-facing wrappers: pointer extraction, stream, argument checks."""
from __future__ import annotations

import torch


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: pats_b200 is CUDA-only (no CPU fallback). "
            "Use pats_b200.host for host buffers."
        )
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def scalar_on(device, value, name: str) -> torch.Tensor:
    """A 1-element f32 device tensor holding `value` (tensor or number) without a host sync."""
    if isinstance(value, torch.Tensor):
        if value.numel() != 1:
            raise ValueError(f"{name} must have exactly one element")
        return value.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.full((1,), float(value), dtype=torch.float32, device=device)
