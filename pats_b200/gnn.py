"""The attention network in front of every matching level on the CUDA library (csrc/gnn.cu).

Mirrors models/modules.py of zju3dv/pats:
    AttentionalGNN.forward(self, desc0, desc1)          :126-134     -> `attentional_gnn_forward` (same signature, bound by
                                                                         pats_b200.install.install(attention=True))
with AttentionalPropagation :108-117, MultiHeadedAttention :90-106, attention :84-88 and MLP :58-69 inside.  The module object
stays the reference's own (its parameters are read, never copied back); the packed form of its weights (query / key / value rows
head-major, merge convolution + inference BatchNorm folded into the first MLP convolution) is built once by the library
(`pats_gnn_pack_f32`) and cached on the module until a parameter changes.

A module in train() mode is this path too (the reference keeps the third layer's network in train() when `if_local` is False,
models/pats.py:112-119 -- three of its four configurations): each BatchNorm call then normalises with the statistics of its batch,
and the module's running buffers and `num_batches_tracked` are updated exactly as the two calls per layer of the reference update
them (`pats_attentional_gnn_train_f32`).  What is NOT this path: head sizes csrc/gnn.cu has no attention kernel for and a BatchNorm
with `momentum=None`; those calls run the reference's own layer modules, exactly as `AttentionalGNN.forward` does -- on the GPU, in
PyTorch; there is no CPU path here either.
"""
from __future__ import annotations

import torch

from . import _lib
from ._torchutil import cuda_f32, stream_ptr

__all__ = ["attentional_gnn_forward", "attentional_gnn", "pack_module", "pack_raw", "supported", "set_precision"]

WORKSPACE_MB = 2048  # activations of one chunk of problems (28 * n * D floats each: FP32 and TF32-half copies); larger chunks measured faster (fewer kernel tails: 300 windows 27.2 ms in chunks of 119, 26.2 ms in one)
_PARAM_ORDER = ("attn.proj.0", "attn.proj.1", "attn.proj.2", "attn.merge", "mlp.0")


_WORKSPACES: dict = {}


def _workspace(device: torch.device, nfloats: int) -> torch.Tensor:
    """One grow-only scratch tensor per (device, stream): a network call needs up to 2 GB, and asking the caching allocator for a block
    of that size inside a forward pass that allocates hundreds of small tensors in between costs a cudaMalloc (and a synchronising
    cudaFree of cached blocks) per call -- measured: 11.7 instead of 13.9 pairs/s through the `if_local=False` forward pass.  Calls on
    one stream are ordered, so consecutive calls may share it."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nfloats:
        _WORKSPACES.pop(key, None)
        ws = torch.empty(int(nfloats), dtype=torch.float32, device=device)
        _WORKSPACES[key] = ws
    return ws


def set_precision(passes: int) -> None:
    """3 (default): FP32-class 3xTF32 convolutions; 1: single-pass TF32 (what cuDNN gives the reference's Conv1d on a GPU)."""
    _lib.load().pats_gnn_precision(int(passes))


def supported(n_tokens: int, d_model: int, heads: int) -> bool:
    """Shapes csrc/gnn.cu has kernels for (pats_attentional_gnn_f32 in include/pats_b200.h)."""
    if d_model % heads or d_model % 8:
        return False
    dim = d_model // heads
    if dim % 2:
        return False
    return dim <= 128  # n <= 96 / dim <= 32 and n <= 160 / dim <= 96: keys resident; anything else: chunked keys, online softmax


def _raw(gnn: torch.nn.Module) -> torch.Tensor:
    """The module's parameters in the order include/pats_b200.h documents for `raw`."""
    parts = []
    for layer in gnn.layers:
        sd = dict(layer.named_parameters())
        sd.update(dict(layer.named_buffers()))
        for k in _PARAM_ORDER:
            parts += [sd[k + ".weight"], sd[k + ".bias"]]
        parts += [sd["mlp.1.weight"], sd["mlp.1.bias"], sd["mlp.1.running_mean"], sd["mlp.1.running_var"], sd["mlp.3.weight"], sd["mlp.3.bias"]]
    return torch.cat([p.detach().reshape(-1).float() for p in parts])


def _slots(gnn: torch.nn.Module):
    """(dict, name) of every parameter / buffer of the network's layers, in module order, cached on the module.  Looking a tensor up
    through its owner's `_parameters` / `_buffers` dict follows replacements (`module.to()` swaps buffer objects, a user may assign a new
    Parameter) at the price of a dict access; walking `layer.parameters()` on every call cost 1.1 ms per call -- 19 ms per image pair."""
    slots = getattr(gnn, "_pats_b200_slots", None)
    if slots is None or slots[0] != len(gnn.layers):
        par, buf = [], []
        for layer in gnn.layers:
            for m in layer.modules():
                par += [(m._parameters, n) for n, p in m._parameters.items() if p is not None]
                buf += [(m._buffers, n) for n, b in m._buffers.items() if b is not None and b.is_floating_point()]
        slots = (len(gnn.layers), par, buf)
        gnn._pats_b200_slots = slots
    return slots


def _key_params(gnn: torch.nn.Module):
    """train(): the packed weights do not depend on the running buffers (which every call updates)"""
    return tuple([(d[n].data_ptr(), d[n]._version) for d, n in _slots(gnn)[1]])


def _key(gnn: torch.nn.Module):
    _, par, buf = _slots(gnn)
    return tuple([(d[n].data_ptr(), d[n]._version) for d, n in par]) + tuple([(d[n].data_ptr(), d[n]._version) for d, n in buf])


def pack_raw(raw: torch.Tensor, layers: int, d_model: int, heads: int, bn_eps: float = 1e-5) -> torch.Tensor:
    """`raw` (CUDA, float32): the parameters of `layers` AttentionalPropagation layers in the order include/pats_b200.h documents
    -> the packed weights `attentional_gnn` takes."""
    lib = _lib.load()
    if not raw.is_cuda:
        raise RuntimeError("pats_b200.gnn: the parameters are on the CPU; pats_b200 is CUDA-only (no CPU fallback)")
    raw = raw.detach().float().contiguous()
    if raw.numel() != lib.pats_gnn_raw_floats(layers, d_model):
        raise RuntimeError(f"pats_b200.gnn: {raw.numel()} parameters given, the packed layout expects {lib.pats_gnn_raw_floats(layers, d_model)}")
    packed = torch.empty(lib.pats_gnn_packed_floats(layers, d_model), dtype=torch.float32, device=raw.device)
    with torch.cuda.device(raw.device):
        rc = lib.pats_gnn_pack_f32(raw.data_ptr(), layers, d_model, heads, float(bn_eps), packed.data_ptr(), stream_ptr(raw.device))
    _lib.check(rc, "gnn_pack")
    return packed


def pack_module(gnn: torch.nn.Module):
    """(packed weights, cross flags, D, heads, layers) of a reference `AttentionalGNN`, cached on the module."""
    key = _key(gnn)
    cached = getattr(gnn, "_pats_b200_pack", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    layer0 = gnn.layers[0]
    D = layer0.attn.merge.weight.shape[0]
    heads = layer0.attn.num_heads
    L = len(gnn.layers)
    packed = pack_raw(_raw(gnn), L, D, heads, float(layer0.mlp[1].eps))
    cross = bytes(1 if n == "cross" else 0 for n in gnn.names)
    out = (packed, cross, D, heads, L)
    gnn._pats_b200_pack = (key, out)
    return out


def attentional_gnn(packed: torch.Tensor, cross: bytes, heads: int, desc0: torch.Tensor, desc1: torch.Tensor, workspace_mb: int | None = None):
    """desc0, desc1 [B,D,N] -> the two updated descriptor sets (fresh tensors)."""
    d0 = cuda_f32(desc0, "desc0")
    d1 = cuda_f32(desc1, "desc1")
    if d0.dim() != 3 or d0.shape != d1.shape:
        raise ValueError(f"attentional_gnn: expected two [b,d,n] tensors of one shape, got {tuple(d0.shape)} and {tuple(d1.shape)}")
    B, D, N = d0.shape
    lib = _lib.load()
    per = lib.pats_gnn_workspace_floats(1, D, N)
    budget = (WORKSPACE_MB if workspace_mb is None else workspace_mb) * (1 << 20) // 4
    chunk = max(1, min(B, budget // per))
    ws = _workspace(d0.device, per * chunk)
    out0, out1 = torch.empty_like(d0), torch.empty_like(d1)
    with torch.cuda.device(d0.device):
        rc = lib.pats_attentional_gnn_f32(d0.data_ptr(), d1.data_ptr(), B, D, N, packed.data_ptr(), cross, len(cross), heads, out0.data_ptr(),
                                          out1.data_ptr(), ws.data_ptr(), per * chunk, stream_ptr(d0.device))
    _lib.check(rc, "attentional_gnn")
    return out0, out1


def _layerwise(self, desc0, desc1):
    """AttentionalGNN.forward as the reference composes it (models/modules.py:126-134), on the module's own layers."""
    for layer, name in zip(self.layers, self.names):
        if name == 'cross':
            src0, src1 = desc1, desc0
        else:
            src0, src1 = desc0, desc1
        delta0, delta1 = layer(desc0, src0), layer(desc1, src1)
        desc0, desc1 = (desc0 + delta0), (desc1 + delta1)
    return desc0, desc1


def _train_forward(self, desc0, desc1):
    """train(): batch-statistics BatchNorm, running buffers updated as the reference's two calls per layer do."""
    lib = _lib.load()
    key = _key_params(self)
    cached = getattr(self, "_pats_b200_pack_train", None)
    if cached is None or cached[0] != key:
        layer0 = self.layers[0]
        D, heads, L = layer0.attn.merge.weight.shape[0], layer0.attn.num_heads, len(self.layers)
        raw = _raw(self)
        if not raw.is_cuda:
            raise RuntimeError("pats_b200.gnn: the module is on the CPU; pats_b200 is CUDA-only (no CPU fallback)")
        packed = torch.empty(lib.pats_gnn_packed_floats(L, D), dtype=torch.float32, device=raw.device)
        with torch.cuda.device(raw.device):
            rc = lib.pats_gnn_pack_train_f32(raw.data_ptr(), L, D, heads, packed.data_ptr(), stream_ptr(raw.device))
        _lib.check(rc, "gnn_pack_train")
        cached = (key, (packed, raw, bytes(1 if n == "cross" else 0 for n in self.names), D, heads, L))
        self._pats_b200_pack_train = cached
    packed, raw, cross, D, heads, L = cached[1]
    d0, d1 = cuda_f32(desc0, "desc0"), cuda_f32(desc1, "desc1")
    B, _, N = d0.shape
    bns = [layer.mlp[1] for layer in self.layers]
    running = torch.stack([torch.stack([bn.running_mean, bn.running_var]) for bn in bns]).float().contiguous()  # [L, 2, 2D]
    need = lib.pats_gnn_workspace_floats(B, D, N)
    ws = _workspace(d0.device, need)
    out0, out1 = torch.empty_like(d0), torch.empty_like(d1)
    with torch.cuda.device(d0.device):
        rc = lib.pats_attentional_gnn_train_f32(d0.data_ptr(), d1.data_ptr(), B, D, N, packed.data_ptr(), raw.data_ptr(), running.data_ptr(),
                                                float(bns[0].momentum), float(bns[0].eps), cross, L, heads, out0.data_ptr(), out1.data_ptr(),
                                                ws.data_ptr(), need, stream_ptr(d0.device))
    _lib.check(rc, "attentional_gnn_train")
    for bn, r in zip(bns, running):  # the side effects of the 2 L BatchNorm calls
        bn.running_mean.copy_(r[0])
        bn.running_var.copy_(r[1])
        bn.num_batches_tracked += 2
    return out0, out1


def attentional_gnn_forward(self, desc0, desc1):
    """AttentionalGNN.forward (models/modules.py:126-134)."""
    if not desc0.is_cuda:
        raise RuntimeError(f"desc0 is on {desc0.device}: pats_b200 is CUDA-only (no CPU fallback)")
    heads = self.layers[0].attn.num_heads
    if desc0.dim() != 3 or not supported(desc0.shape[2], desc0.shape[1], heads):
        return _layerwise(self, desc0, desc1)
    if self.training:
        bn = self.layers[0].mlp[1]
        if bn.momentum is None or not bn.track_running_stats or desc0.shape[0] * desc0.shape[2] < 2:
            return _layerwise(self, desc0, desc1)
        return _train_forward(self, desc0, desc1)
    packed, cross, _, heads, _ = pack_module(self)
    return attentional_gnn(packed, cross, heads, desc0, desc1)
