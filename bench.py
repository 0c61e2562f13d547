#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the PATS OT + subdivision hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs-per-step B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A *step* is one pass of the hot path over a batch of `pairs_per_step` synthetic 640x480 image pairs, with the
shapes the reference produces for such a pair:
    level 1  log_optimal_transport  1 x (300 -> 301)^2, 100 it  -> est_position (15x20 grid) -> Compute_imgs (300 patches)
    level 2  log_optimal_transport2 300 x 145^2, 100 it         -> est_position (12x12 grid) -> merge_patches_new
    level 3  log_optimal_transport2 4800 x 65^2, 100 it         -> Compute_result           -> get_result
(K = 4800 = every 8-px cell of the 60x80 fine grid owned by exactly one window after the merge.)
The ResNet / attention layers between the stages are out of scope (host PyTorch in the reference), so each
stage's network-produced inputs (scores, scales) are synthetic, seeded, and independent; inside a stage the data
flows on the device exactly as in the reference (OT plan -> est_position / Compute_result; bounds -> patches).

`value`   : pairs/s with every input already resident in HBM, kernels launched through the C ABI on one stream.
`e2e`     : pairs/s through the public torch-facing API with HOST (pinned) inputs: per step the stage inputs are
            copied host->device, the reference-named wrappers run, and the results (matches, masks, fine points)
            are read back device->host, all inside the timed region.
`roofline`: the dominant kernel (level-3 Sinkhorn, one warp per 65x65 problem), timed with CUDA events on the
            launching stream inside the timed region.
`cpu_baseline` / --impl reference: the CPU oracle port (+ the compiled reference tensor_resize, oracle/_ref) on the
            box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "image-pairs/sec (640x480, 100 Sinkhorn it)"
H, W_IMG, PS = 480, 640, 32
GH, GW = H // PS, W_IMG // PS          # 15 x 20 coarse patches
N1 = GH * GW                           # 300
P2 = 300                               # matched windows of one pair (all coarse patches matched)
K3 = 4800                              # level-3 problems of one pair (60 x 80 fine cells)
ITERS = 100
SEED = 18027                           # configs/*.yaml `seed`
DEFAULT_PAIRS_PER_STEP = 8             # pairs are independent (evaluate.py:25-35): a step batches 8 of them through every kernel
LOG2_F32 = 0.6931471824645996          # f32(log(f32(2))): torch.log(self.one * 2), second_layer.py:109-110 (outdoor)
L3_SAMPLE_EVERY = 8                    # every 8th timed step brackets the level-3 solve with events (roofline sample)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torch warnings) also write to fd 1, so the
# real stdout is saved here and fd 1 is pointed at stderr for the rest of the run; emit() writes the line to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs of one step (seeded; host tensors)
# ------------------------------------------------------------------------------------------------------------
WORKLOAD = "pats_hot_path_pair640x480"
SURVIVE = 0.01                         # fraction of the 691 200 fine points of a pair that end up in its match list: ~6.9 k matches per pair
                                       # (the live conditioned forward keeps 0.4 - 1.7 k, tests/test_gpu_live_forward.py; a trained PATS a few thousand)


def planted_scores(torch, g, b, gh, gw, sharp, noise, floor, dustbin=False, peak=5.0):
    """Affinity of a smooth random warp between two gh x gw grids: peaked, spatially coherent plans, i.e. what a trained matcher
    hands to the Sinkhorn (and what the conditioned live forward produces: |0.1 * scores| <= 12 .. 26, std 2.5).  With random
    `0.1 * randn` scores -- the `diffuse` workload -- every plan is flat, boxes grow to the iteration limit and the log-domain
    paths are never touched.  Returns (scores, area [b]): area = source cells per target cell of the warp, the consistent value of the
    target areas `ns` (first_layer.py:106-107).  dustbin=True appends the dustbin row / column of log_optimal_transport2's [b, n+1, n+1] input."""
    n = gh * gw
    ys, xs = torch.meshgrid(torch.arange(gh).float(), torch.arange(gw).float(), indexing="ij")
    src = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
    ctr = torch.tensor([gh / 2.0, gw / 2.0])
    out = torch.empty(b, n + int(dustbin), n + int(dustbin))
    area = torch.empty(b)
    step = max(1, min(b, (1 << 24) // (n * n)))
    for lo in range(0, b, step):
        c = min(step, b - lo)
        A = torch.eye(2)[None] * (0.6 + 0.8 * torch.rand(c, 1, 1, generator=g)) + 0.1 * torch.randn(c, 2, 2, generator=g)
        t = torch.randn(c, 1, 2, generator=g) * (0.12 * max(gh, gw))  # part of the source cells land outside the target grid: they feed the dustbin
        area[lo:lo + c] = (1.0 / torch.linalg.det(A).abs()).clamp(1.0 / 16, 16.0)  # source cells per target cell: the target "area" ns
        warped = (src - ctr) @ A.transpose(1, 2) + ctr + t
        d2 = ((warped[:, :, None, :] - src[None, None, :, :]) ** 2).sum(-1)
        sc = (peak - d2 / sharp + noise * torch.randn(c, n, n, generator=g)).clamp_min(floor)  # a true match scores ~peak: it must beat the dustbin (score ~ -1 .. 1, mass = n) by ln(n) + a few nats to keep its row
        if dustbin:
            out[lo:lo + c, :n, :n] = sc
            out[lo:lo + c, n, :] = -1.0 + 0.3 * torch.randn(c, n + 1, generator=g)
            out[lo:lo + c, :, n] = -1.0 + 0.3 * torch.randn(c, n + 1, generator=g)
        else:
            out[lo:lo + c] = sc
    return out, area


def make_inputs(torch, pairs: int, seed: int, kind: str = "planted"):
    g = torch.Generator().manual_seed(seed)

    def areas(*shape, span):
        return torch.exp((torch.rand(*shape, generator=g) * 2 - 1) * math.log(span))

    B = pairs
    d = {}
    if kind == "planted":
        d["l1_scores"], a1 = planted_scores(torch, g, B, GH, GW, sharp=1.5, noise=0.3, floor=-20.0, peak=12.0)      # first_layer.py:110-114 (already x0.1)
        d["l1_ns"] = (a1.reshape(B, 1, 1) * areas(B, 1, N1, span=1.2)).contiguous()                       # first_layer.py:106-107
    else:
        d["l1_scores"] = 0.1 * torch.randn(B, N1, N1, generator=g)
        d["l1_ns"] = areas(B, 1, N1, span=16.0)
    d["alpha"] = torch.tensor([1.0])                                      # |bin_score| (init 0.0 would switch the dustbin off)
    d["left"] = torch.randint(0, 256, (B, H, W_IMG, 3), generator=g, dtype=torch.uint8)   # evaluate.py:26-27
    d["right"] = torch.roll(d["left"], (16, 24), dims=(1, 2)).contiguous()
    # Compute_imgs inputs for the `diffuse` workload, whose est_position output is meaningless (the planted workload feeds
    # Compute_imgs from est_position on the device, as first_layer.py:121-140 does)
    d["ci_xs"] = areas(B, N1, span=1.6)
    d["ci_ys"] = areas(B, N1, span=1.6)
    cy = torch.arange(GH).float().reshape(1, GH, 1).expand(B, GH, GW) + 0.5 + 0.6 * torch.randn(B, GH, GW, generator=g)
    cx = torch.arange(GW).float().reshape(1, 1, GW).expand(B, GH, GW) + 0.5 + 0.6 * torch.randn(B, GH, GW, generator=g)
    d["ci_avg"] = torch.stack([cy, cx], -1).reshape(B, N1, 2).contiguous()
    d["ci_nm"] = torch.zeros(B, N1, dtype=torch.bool)
    if kind == "planted":
        d["l2_scores"], a2 = planted_scores(torch, g, B * P2, 12, 12, sharp=1.5, noise=0.3, floor=-20.0, dustbin=True, peak=9.0)   # second_layer.py:100-104
        d["l2_sx"] = (a2.sqrt().reshape(-1, 1) * areas(B * P2, 144, span=1.1)).contiguous()                               # second_layer.py:92-98
        d["l2_sy"] = (a2.sqrt().reshape(-1, 1) * areas(B * P2, 144, span=1.1)).contiguous()
    else:
        d["l2_scores"] = 0.1 * torch.randn(B * P2, 145, 145, generator=g)
        d["l2_sx"] = areas(B * P2, 144, span=16.0)
        d["l2_sy"] = areas(B * P2, 144, span=16.0)
    d["l2_ns"] = (d["l2_sx"] * d["l2_sy"]).reshape(B * P2, 1, 144).contiguous()
    d["one"] = torch.tensor([1.0])
    d["nm1_L1"] = torch.zeros(B, N1, dtype=torch.bool)
    if kind == "planted":
        d["l3_scores"], a3 = planted_scores(torch, g, B * K3, 8, 8, sharp=1.5, noise=0.3, floor=-10.0, dustbin=True, peak=7.0)     # third_layer.py:156-158
        d["l3_ns"] = (a3.reshape(-1, 1, 1) * areas(B * K3, 1, 64, span=1.2)).contiguous()                                 # third_layer.py:151-152
    else:
        d["l3_scores"] = 0.1 * torch.randn(B * K3, 65, 65, generator=g)
        d["l3_ns"] = areas(B * K3, 1, 64, span=16.0)
    d["l3_sxy"] = (d["l3_ns"].reshape(B * K3, 64) + 1e-8).sqrt()           # third_layer.py:153-154
    d["p_s"] = torch.randint(0, 24, (B * K3, 2), generator=g) * 4
    d["p_t"] = torch.randint(0, 25, (B * K3, 2), generator=g) * 4
    # get_result inputs (models/pats.py:72-78): level-0 geometry per window, level-1 masks / points per fine cell
    d["gr_nm0"] = torch.zeros(B, N1, dtype=torch.bool)
    d["gr_pt0"] = (d["ci_avg"].flip(2) / 1.0).contiguous()
    d["gr_sc0"] = torch.cat([areas(B, N1, 1, span=1.6), torch.ones(B, N1, 1)], 2).contiguous()
    d["gr_nm1"] = torch.rand(B * P2, 2304, generator=g) >= (SURVIVE if kind == "planted" else 0.25)
    d["gr_pt1"] = torch.rand(B * P2, 2304, 2, generator=g) * 48
    d["gr_sc1"] = d["gr_sc0"].reshape(B * N1, 1, 2).repeat(1, 2304, 1).contiguous()
    return d


# ------------------------------------------------------------------------------------------------------------
# device-resident step through the C ABI (no allocation, no host sync inside)
# ------------------------------------------------------------------------------------------------------------
class DeviceStep:
    LAUNCHES = 0  # kernels of ours enqueued per step (counted from the launch table below)

    def __init__(self, torch, dev, pairs: int, seed: int, kind: str = "planted"):
        from pats_b200 import _lib

        self.torch, self.dev, self.B = torch, dev, pairs
        self.lib = _lib.load()
        self.check = _lib.check
        self.kind = kind
        host = make_inputs(torch, pairs, seed, kind)
        self.i = {k: v.to(dev) for k, v in host.items()}
        self.host_inputs = host
        B = pairs
        f32, i64, u8, f64 = torch.float32, torch.int64, torch.uint8, torch.float64
        E = lambda *s, dt=f32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        o = {}
        o["l1_Z"] = E(B, N1 + 1, N1 + 1)
        for lvl, b, n in (("l1", B, N1), ("l2", B * P2, 144)):
            o[lvl + "_trust"], o[lvl + "_xs"], o[lvl + "_ys"], o[lvl + "_core"] = E(b, n), E(b, n), E(b, n), E(b, n)
            o[lvl + "_avg"] = E(b, n, 2)
            o[lvl + "_nm1"], o[lvl + "_nm2"] = E(b, n, dt=u8), E(b, n, dt=u8)
            o[lvl + "_bound"] = E(b, n, 4, dt=i64)
        o["new_left"] = E(B * N1, 96, 96, 3, dt=u8)
        o["new_right"] = E(B * N1, 3, 96, 96)
        o["bound5"] = E(B * N1, 5, dt=i64)
        o["ci_xs_new"], o["ci_ys_new"], o["ci_avg_new"] = E(B, N1, 2), E(B, N1, 2), E(B, N1, 2)
        o["ci_meta"] = torch.zeros(2, dtype=torch.int32, device=dev)
        o["l2_Z"] = E(B * P2, 145, 145)
        o["scores_back"] = torch.zeros(B, N1, 16, 9, dtype=f64, device=dev)
        o["merge_out"] = E(B * P2, 144, dt=u8)
        o["merge_ws"] = torch.empty(2 * B * N1 + 1, dtype=torch.int32, device=dev)
        o["l3_Z"] = E(B * K3, 65, 65)
        o["mk0"], o["mk1"] = E(B * K3, 16, 2), E(B * K3, 16, 2)
        o["im1"] = E(B * K3, 16, dt=u8)
        cap = B * P2 * 2304
        o["ml"], o["mr"] = E(cap, 2), E(cap, 2)
        o["gr_total"] = torch.zeros(1, dtype=i64, device=dev)
        o["gr_ws"] = torch.empty(8 * (B * P2 + 1) + 4 * (2 * B * N1 + 1 + B * P2) + 8, dtype=u8, device=dev)
        self.o = o
        self.gr_nm1_u8 = self.i["gr_nm1"].to(u8)
        self.ev_l3 = []

    def run(self, stream_ptr: int, time_l3=None):
        L, i, o, B, c = self.lib, self.i, self.o, self.B, self.check
        p = lambda t: t.data_ptr()  # noqa: E731
        n = 0
        # ---- level 1 ----------------------------------------------------------------------------------------
        c(L.pats_log_optimal_transport_f32(p(i["l1_scores"]), p(i["alpha"]), p(i["l1_ns"]), B, N1, N1, ITERS, p(o["l1_Z"]), stream_ptr), "ot1"); n += 1
        c(L.pats_est_position_f32(p(o["l1_Z"]), p(i["l1_ns"]), p(i["l1_ns"]), B, GH, GW, 1e-5, 15, p(o["l1_trust"]), p(o["l1_avg"]), p(o["l1_xs"]),
                                  p(o["l1_ys"]), p(o["l1_nm1"]), p(o["l1_nm2"]), p(o["l1_core"]), p(o["l1_bound"]), stream_ptr), "est1"); n += 1
        # first_layer.py:121-140: est_position's areas, centres and mask ARE Compute_imgs' arguments (planted workload); the diffuse
        # workload's est_position output is noise, so there Compute_imgs keeps its own seeded inputs
        chain = self.kind == "planted"
        c(L.pats_compute_imgs(p(o["l1_xs"] if chain else i["ci_xs"]), p(o["l1_ys"] if chain else i["ci_ys"]), p(o["l1_avg"] if chain else i["ci_avg"]),
                              p(o["l1_nm1"] if chain else i["ci_nm"]), p(i["left"]), p(i["right"]), 1, B, GH, GW, PS, 128,
                              p(o["new_left"]), p(o["new_right"]), p(o["bound5"]), p(o["ci_xs_new"]), p(o["ci_ys_new"]), p(o["ci_avg_new"]), B * N1,
                              p(o["ci_meta"]), p(o["ci_meta"]) + 4, stream_ptr), "imgs"); n += 3
        # ---- level 2 (second_layer.py:103-116 in one call: OT -> dustbin offsets -> est_position, handed over per problem) ----
        c(L.pats_second_layer_match_f32(p(i["l2_scores"]), p(i["one"]), p(i["l2_ns"]), p(i["l2_sx"]), p(i["l2_sy"]), B * P2, 12, 12, ITERS, LOG2_F32,
                                        1e-3, 8, p(o["l2_Z"]), p(o["l2_trust"]), p(o["l2_avg"]), p(o["l2_xs"]), p(o["l2_ys"]), p(o["l2_nm1"]),
                                        p(o["l2_nm2"]), p(o["l2_core"]), p(o["l2_bound"]), stream_ptr), "second_layer_match"); n += 2
        c(L.pats_merge_patches(1, p(o["l2_trust"]), p(i["nm1_L1"]), p(o["l2_nm1"]), p(o["scores_back"]), B, GH, GW, B * P2, p(o["merge_out"]),
                               p(o["merge_ws"]), stream_ptr), "merge"); n += 3
        # ---- level 3 (third_layer.py:158-167 in one call: OT -> exp -> Compute_result + label test) ---------------------
        if time_l3 is not None:
            # the roofline sample: the solve alone between two events (an event between the two kernels would serialise the
            # hand-over, so the sampled steps run the two calls separately; they stay inside the timed region)
            time_l3[0].record()
            c(L.pats_log_optimal_transport2_f32(p(i["l3_scores"]), p(i["one"]), p(i["l3_ns"]), B * K3, 65, 65, ITERS, p(o["l3_Z"]), stream_ptr), "ot3")
            time_l3[1].record()
            c(L.pats_third_result_from_log_f32(p(o["l3_Z"]), p(i["l3_sxy"]), p(i["l3_sxy"]), p(i["p_s"]), p(i["p_t"]), B * K3, p(o["mk0"]), p(o["mk1"]),
                                               p(o["im1"]), stream_ptr), "third")
        else:
            c(L.pats_third_layer_match_f32(p(i["l3_scores"]), p(i["one"]), p(i["l3_ns"]), p(i["l3_sxy"]), p(i["l3_sxy"]), p(i["p_s"]), p(i["p_t"]),
                                           B * K3, ITERS, p(o["l3_Z"]), p(o["mk0"]), p(o["mk1"]), p(o["im1"]), stream_ptr), "third_layer_match")
        n += 2
        c(L.pats_get_result_f32(p(i["gr_nm0"]), p(i["gr_pt0"]), p(i["gr_sc0"]), B, 32, GH, GW, p(self.gr_nm1_u8), p(i["gr_pt1"]), p(i["gr_sc1"]),
                                B * P2, 2, 48, 48, p(o["ml"]), p(o["mr"]), B * P2 * 2304, p(o["gr_total"]), p(o["gr_ws"]), stream_ptr), "result")
        n += 3 if B * P2 <= 2048 else 4  # window_maps, count_rows, (exclusive_scan for many windows), assemble_matches
        DeviceStep.LAUNCHES = n
        return n


# ------------------------------------------------------------------------------------------------------------
# end-to-end step through the torch-facing public API with host (pinned) buffers
# ------------------------------------------------------------------------------------------------------------
_WC_KEEP = []


def wc_pinned(torch, nbytes: int):
    """A write-combined pinned host buffer as a uint8 tensor (cudaHostAlloc + cudaHostAllocWriteCombined), or None.  The CPU only
    ever WRITES the staging arena; write-combined pages are not snooped during the device's reads over PCIe, which matters when
    eight ranks pull from one host memory system at once (A/B: --wc-staging)."""
    try:
        import ctypes as C

        rt = C.CDLL("libcudart.so.12")
        ptr = C.c_void_p()
        if rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(nbytes), C.c_uint(0x04)) != 0 or not ptr.value:
            return None
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        t = torch.frombuffer(buf, dtype=torch.uint8)
        _WC_KEEP.append((rt, ptr, buf))
        return t if t.is_pinned() else None
    except Exception:  # noqa: BLE001
        return None


class E2EStep:
    """One step = H2D of every stage input from pinned host memory, the reference-named wrappers, D2H of the results.
    Steps are software-pipelined over two device input buffers: the copies of step i+1 run on a copy stream while
    step i computes (every step still pays its own H2D and D2H inside the timed region)."""

    def __init__(self, torch, dev, pairs: int, host_inputs, kind: str = "planted", wc: bool = False):
        self.torch, self.dev, self.B, self.kind = torch, dev, pairs, kind
        # every stage input lives in ONE pinned host arena and goes over in ONE cudaMemcpyAsync per step (a copy per
        # tensor costs ~17 DMA set-ups per step); the device-side tensors are views into the arena's device twin
        offs, total = {}, 0
        for k, v in host_inputs.items():
            total = (total + 255) & ~255
            offs[k] = total
            total += v.numel() * v.element_size()
        self.h_arena = wc_pinned(torch, total) if wc else None
        self.staging = "write-combined pinned (cudaHostAllocWriteCombined)" if self.h_arena is not None else "pinned"
        if self.h_arena is None:
            self.h_arena = torch.empty(total, dtype=torch.uint8).pin_memory()
        for k, v in host_inputs.items():
            self.h_arena[offs[k]:offs[k] + v.numel() * v.element_size()].view(v.dtype).reshape(v.shape).copy_(v)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host_inputs.values())
        self.d_arena = [torch.empty(total, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.bufs = [{k: ar[offs[k]:offs[k] + v.numel() * v.element_size()].view(v.dtype).reshape(v.shape) for k, v in host_inputs.items()}
                     for ar in self.d_arena]
        self.d2h_bytes = 0
        self.out_host = None
        self.out_pinned = [None, None]
        self.copy_stream = torch.cuda.Stream(dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]

    def _enqueue_copy(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            self.d_arena[slot].copy_(self.h_arena, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def _compute(self, slot):
        torch, dev, B = self.torch, self.dev, self.B
        from pats_b200 import layers as Ly
        from pats_b200 import modules as M
        from pats_b200 import utils as U

        cur = torch.cuda.current_stream(dev)
        cur.wait_event(self.ready[slot])
        d = self.bufs[slot]
        # level 1
        Z1 = M.log_optimal_transport(d["l1_scores"], d["alpha"], d["l1_ns"], ITERS)
        trust1, avg1, xs1, ys1, nm1a, nm1b = Ly.est_position(Z1, d["l1_ns"], d["l1_ns"], GH, GW, 15, 1e-5)
        if self.kind == "planted":  # first_layer.py:121-140: est_position's outputs are Compute_imgs' arguments
            new_left, new_right, xsn, ysn, avn = U.Compute_imgs(xs1, ys1, avg1, nm1a, d["left"], d["right"], width=GW, height=GH)
        else:
            new_left, new_right, xsn, ysn, avn = U.Compute_imgs(d["ci_xs"], d["ci_ys"], d["ci_avg"], d["ci_nm"], d["left"], d["right"], width=GW, height=GH)
        # level 2
        Z2, trust2, avg2, xs2, ys2, nm2a, nm2b = Ly.second_layer_match(d["l2_scores"], 1.0, d["l2_ns"], d["l2_sx"], d["l2_sy"], ITERS, True, 12)
        sb = torch.zeros(B, N1, 16, 9, dtype=torch.float64, device=dev)
        keep, sb = Ly.merge_patches_new(None, B * P2, trust2, [H, W_IMG], d["nm1_L1"], nm2a, sb)
        # level 3
        Z3, mk0, mk1, im1 = Ly.third_layer_match(d["l3_scores"], 1.0, d["l3_ns"], d["l3_sxy"], d["l3_sxy"], d["p_s"], d["p_t"], ITERS)
        ml, mr = U.get_result(B, [d["gr_nm0"], d["gr_nm1"]], [d["gr_pt0"], d["gr_pt1"]], [d["gr_sc0"], d["gr_sc1"]], [[32, GH, GW], [2, 48, 48]], None)
        # D2H of the step's results: the match lists and what the next (out-of-scope) network stages / the caller consume.
        # Pinned destinations, asynchronous copies, one event wait (get_result already synchronised once for the match count).
        outs = [ml, mr, mk0, mk1, im1, keep, trust1, avg1, xs1, ys1, nm1a, nm1b, avg2, xsn, ysn, avn]
        pins = self.out_pinned[slot]
        if pins is None or any(p.shape != t.shape for p, t in zip(pins, outs)):
            pins = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs]
            self.out_pinned[slot] = pins
        for p_, t in zip(pins, outs):
            p_.copy_(t, non_blocking=True)
        self.free[slot].record(cur)
        self.done[slot].record(cur)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in outs)
        return slot

    def _collect(self, slot):
        self.done[slot].synchronize()
        self.out_host = self.out_pinned[slot]
        return self.out_host

    def run(self, n_steps: int = 1):
        self.torch.cuda.synchronize(self.dev)
        for ev in self.free:
            ev.record(self.torch.cuda.current_stream(self.dev))
        self._enqueue_copy(0)
        out, pending = None, None
        for i in range(n_steps):
            if i + 1 < n_steps:
                self._enqueue_copy((i + 1) & 1)
            slot = self._compute(i & 1)
            if pending is not None:  # step i-1's results are complete on the host while step i runs (its pinned buffers are
                out = self._collect(pending)  # not written again before step i+1)
            pending = slot
        return self._collect(pending)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port (+ compiled reference tensor_resize) on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------
def cpu_sample_step(torch, host_inputs, frac_l3: float, frac_l2: float):
    """Runs the CPU implementation of one pair's hot path on a sample; returns (seconds scaled to ONE full pair,
    description).  Level-2 / level-3 problems are independent, so their time scales linearly with the count."""
    import numpy as np

    import oracle

    d = {k: v.numpy() for k, v in host_inputs.items()}
    n3 = max(1, int(K3 * frac_l3))
    n2 = max(1, int(P2 * frac_l2))
    t = {}
    t0 = time.perf_counter()
    Z1 = oracle.log_optimal_transport(d["l1_scores"][:1], 1.0, d["l1_ns"][:1], ITERS)
    t["ot1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.est_nomatching(Z1, N1)
    oracle.iterative_expand_matrix(np.exp(Z1), d["l1_ns"][:1].reshape(1, N1), d["l1_ns"][:1].reshape(1, N1), GH, GW, 1e-5, 15)
    t["est1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    try:  # the reference's own native op where it was compiled (oracle/_ref), else the oracle port
        from oracle import build_ref

        ref_mod = build_ref.load_ref()
        bound, xs, ys, avg = oracle.compute_bounds(d["ci_xs"][:1], d["ci_ys"][:1], d["ci_avg"][:1], GH, GW)
        seq = np.arange(N1).reshape(-1, 1)
        b5 = torch.from_numpy(np.concatenate([bound[0], seq], 1))
        right_use = torch.nn.functional.pad(torch.from_numpy(d["right"][:1]), (0, 0, 128, 128, 128, 128)).permute(0, 3, 1, 2).float()
        ref_mod.tensor_resize(right_use, b5)
        left_use = np.ascontiguousarray(np.pad(d["left"][:1], ((0, 0), (32, 32), (32, 32), (0, 0))).transpose(0, 3, 1, 2))
        oracle.origin_extract(left_use, 32, GW, GH)
        resize_kind = "reference library.cpp (oracle/_ref)"
    except Exception:
        oracle.compute_imgs(d["ci_xs"][:1], d["ci_ys"][:1], d["ci_avg"][:1], d["ci_nm"][:1], d["left"][:1], d["right"][:1], width=GW, height=GH)
        resize_kind = "oracle port"
    t["imgs"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Z2 = oracle.log_optimal_transport2(d["l2_scores"][:n2], 1.0, d["l2_ns"][:n2], ITERS)
    t["ot2"] = (time.perf_counter() - t0) * (P2 / n2)
    t0 = time.perf_counter()
    oracle.est_nomatching(Z2, 144)
    w2 = oracle.iterative_expand_matrix(np.exp(Z2), d["l2_sx"][:n2], d["l2_sy"][:n2], 12, 12, 1e-3, 8)
    t["est2"] = (time.perf_counter() - t0) * (P2 / n2)
    t0 = time.perf_counter()
    trust_full = np.tile(w2[0], (math.ceil(P2 / n2), 1))[:P2]
    nm_full = np.tile(w2[6], (math.ceil(P2 / n2), 1))[:P2]
    oracle.merge_patches(True, trust_full, [H, W_IMG], d["nm1_L1"][:1], nm_full, np.zeros((1, N1, 16, 9)))
    t["merge"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Z3 = oracle.log_optimal_transport2(d["l3_scores"][:n3], 1.0, d["l3_ns"][:n3], ITERS)
    t["ot3"] = (time.perf_counter() - t0) * (K3 / n3)
    t0 = time.perf_counter()
    oracle.third_compute_result(np.exp(Z3), d["l3_sxy"][:n3], d["l3_sxy"][:n3], d["p_s"][:n3], d["p_t"][:n3])
    t["third"] = (time.perf_counter() - t0) * (K3 / n3)
    t0 = time.perf_counter()
    oracle.get_result([d["gr_nm0"][:1], d["gr_nm1"][:P2]], [d["gr_pt0"][:1], d["gr_pt1"][:P2]], [d["gr_sc0"][:1], d["gr_sc1"][:P2]],
                      [[32, GH, GW], [2, 48, 48]])
    t["result"] = time.perf_counter() - t0
    if n2 == P2 and n3 == K3:
        desc = (f"one full 640x480 pair, nothing scaled: L1 OT + est_position + Compute_imgs ({resize_kind}) + all {P2} level-2 problems "
                f"(OT + est_position) + merge + all {K3} level-3 problems (OT + Compute_result) + get_result")
    else:
        desc = (f"one 640x480 pair: L1 OT + est_position + Compute_imgs ({resize_kind}) + merge + get_result in full; "
                f"{n2}/{P2} level-2 and {n3}/{K3} level-3 problems (OT + est_position / Compute_result), time scaled linearly")
    return sum(t.values()), desc, t


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every ~5 ms during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.stop_flag, self.thread, self.err = gpu_index, [], False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def pump():
                while not self.stop_flag:
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception as e:  # noqa: BLE001
                        self.err = repr(e)
                        return
                    time.sleep(0.004)

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(v for k, v in names.items() if bits & k),
                "samples": len(rows)}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak_bf16():
    """dense BF16 TFLOP/s measured on this pool's B200s (MEASURED_PEAKS.json, burst figure), else the recipe's fallback"""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops"])
    except Exception:  # noqa: BLE001
        return 2250.0


def ncu_traffic(key="sinkhorn_warp_dram_bytes_per_launch"):
    path = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(key)
        except Exception:
            return None
    return None


def forward_leg(torch, dev, args, if_local=True, arms=None):
    """image-pairs/s through the UNMODIFIED reference `PATS.forward` (models/pats.py:18-85, driven like evaluate.py:25-33: uint8
    images host -> device, forward, match lists device -> host), three ways on the same seeded random-init network
    (tests/live_util: conditioned to realistic score magnitudes) and the same synthetic 640x480 pairs:
        reference_cuda   the stock reference on CUDA tensors (ATen ops + its compiled setup/library.cpp) -- how it really runs
        installed        + pats_b200.install.install(): every hot-path function on the CUDA library
        installed_fused  + install(fused=True): the two layer forwards on the fused entry points as well
        installed_fused_attention  + install(fused=True, attention=True): the attention networks on csrc/gnn.cu too
        reference_cpu    the stock reference on CPU tensors, all host threads (one pair; --no-forward-cpu skips it)
    The ResNet / attention networks are NOT part of the path (they stay the reference's own PyTorch modules in every arm), so
    this leg bounds what the hot path is worth inside the whole forward pass; `value` / `e2e` above isolate the path itself."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import live_util as L

    if L.reference_root() is None:
        return {"unavailable": "reference Python not staged (oracle/_ref/py: run __graft_entry__.build() where /root/reference exists)"}
    ref = L.load_reference()
    import pats_b200.install as inst

    cfg = L.config(if_local=if_local, merge_new=True, if_outdoor=True)  # configs/test_megadepth.yaml (if_local) / test_yfcc.yaml, test_demo.yaml
    n_pairs, n_warm = (6, 2) if if_local else (5, 2)
    host = [tuple(t.pin_memory() for t in L.synthetic_pair((H, W_IMG), seed=SEED + i)) for i in range(n_pairs)]
    out = {"config": ("configs/test_megadepth.yaml flags (if_local, merge_new, if_outdoor)" if if_local else
                      "configs/test_yfcc.yaml / test_demo.yaml flags (if_local False: no chunking, the third layer's network in train() mode; merge_new, if_outdoor)")
                     + ", 640x480, seeded random-init weights conditioned by tests/live_util.py",
           "pairs_timed": n_pairs - n_warm, "unit": "pairs/s"}

    def run(model, device, pairs):
        n_match = 0
        for i0, i1 in pairs:
            data = {"image0": i0.to(device, non_blocking=True), "image1": i1.to(device, non_blocking=True)}
            r = model(data)
            n_match += int(r["matches_l"].cpu().shape[0]) + 0 * int(r["matches_r"].cpu().shape[0])
        return n_match

    with torch.no_grad():
        model = L.build_model(ref, cfg, device=dev)
        for name, setup in (("reference_cuda", None), ("installed", dict(fused=False)), ("installed_fused", dict(fused=True)),
                            ("installed_fused_attention", dict(fused=True, attention=True))):
            if arms is not None and name not in arms:
                continue
            if setup is not None:
                inst.install(**setup)
            try:
                run(model, dev, host[:n_warm])
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                m = run(model, dev, host[n_warm:])
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
            finally:
                if setup is not None:
                    inst.uninstall()
            out[name] = {"value": (n_pairs - n_warm) / dt, "s_per_pair": dt / (n_pairs - n_warm), "matches": m}
        if not args.no_forward_cpu and arms is None:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cpu_model = L.build_model(ref, cfg, device="cpu")
            orig_cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda self, *a, **k: self  # models/pats.py:76 hard-codes .cuda(); this arm runs on CPU tensors
            try:
                t0 = time.perf_counter()
                m = run(cpu_model, "cpu", host[n_warm:n_warm + 1])
                dt = time.perf_counter() - t0
            finally:
                torch.Tensor.cuda = orig_cuda
            out["reference_cpu"] = {"value": 1.0 / dt, "s_per_pair": dt, "matches": m, "cores": cores, "pairs_timed": 1}
    for arm, key in (("installed", "speedup_installed_vs_reference_cuda"), ("installed_fused", "speedup_fused_vs_reference_cuda"),
                     ("installed_fused_attention", "speedup_fused_attention_vs_reference_cuda")):
        if arm in out and "reference_cuda" in out:
            out[key] = out[arm]["value"] / out["reference_cuda"]["value"]
    return out


def forward_sharded_leg(torch, dist, dev, rank, world, pairs_per_rank=4):
    """SURVEY.md 8(d) / BASELINE configs[3] at driver size: image-pairs/s through the UNMODIFIED `PATS.forward` with the whole path
    installed (install(fused=True, attention=True)) on N GPUs.  Pairs are independent (evaluate.py:25-35): rank r runs its contiguous
    shard of the seeded synthetic pair list (pats_b200.dist.shard_range; images host -> device inside the timed region), keeps the
    match lists on the device, and after the pair loop every list is gathered to rank 0 (dist.gather_match_lists: 2 collectives,
    1 host sync).  Time = barrier .. barrier, max over ranks; weights replicated (same seed on every rank)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import live_util as L

    if L.reference_root() is None:
        return {"unavailable": "reference Python not staged (oracle/_ref/py)"}
    ref = L.load_reference()
    import pats_b200.install as inst
    from pats_b200 import dist as pdist

    cfg = L.config(if_local=True, merge_new=True, if_outdoor=True)  # configs/test_megadepth.yaml
    n_total, n_warm = pairs_per_rank * world, 2
    lo, hi = pdist.shard_range(n_total, rank, world)
    host = {i: tuple(t.pin_memory() for t in L.synthetic_pair((H, W_IMG), seed=SEED + i)) for i in list(range(lo, hi)) + [n_total + k for k in range(n_warm)]}

    def one(i):
        i0, i1 = host[i]
        r = model({"image0": i0.to(dev, non_blocking=True), "image1": i1.to(dev, non_blocking=True)})
        return torch.cat([r["matches_l"].float(), r["matches_r"].float()], 1)

    with torch.no_grad():
        model = L.build_model(ref, cfg, device=dev)
        inst.install(fused=True, attention=True)
        try:
            for k in range(n_warm):
                one(n_total + k)
            torch.cuda.synchronize(dev)
            dist.barrier(device_ids=[dev.index])
            t0 = time.perf_counter()
            lists = [one(i) for i in range(lo, hi)]
            stats = {}
            gathered = pdist.gather_match_lists(lists, max_pairs=pairs_per_rank, dst=0, stats=stats)
            torch.cuda.synchronize(dev)
            dist.barrier(device_ids=[dev.index])
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        finally:
            inst.uninstall()
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out = {"config": "configs/test_megadepth.yaml flags, 640x480, install(fused=True, attention=True), conditioned random-init weights (tests/live_util.py)",
           "pairs": n_total, "pairs_per_rank": pairs_per_rank, "seconds": float(dt.item()), "value": n_total / float(dt.item()), "unit": "pairs/s",
           "exchange": stats}
    if rank == 0:
        out["matches_gathered"] = int(sum(m.shape[0] for per_rank in gathered for m in per_rank))
    return out


def tail_leg(torch, dev, n_pairs=6):
    """N4: the evaluation loop of evaluate.py:20-39 -- dataset item (image decode / resize stand-in: tests/pose_util.py), PATS.forward with
    the whole path installed, the reference's compute_pose_error (utils/metrics.py:21-66: OpenCV essential-matrix RANSAC + recoverPose)
    -- run stage after stage as the reference does, and as pats_b200.pipeline.evaluate_pairs' three-stage pipeline (same calls, same
    order, bit-identical pose errors: tests/test_gpu_pipeline.py).  The forward pass runs on the real (random-init, conditioned)
    network; RANSAC is fed planted two-view correspondences with 30 % outliers (random-init matches fit no essential matrix and would
    make RANSAC run to its iteration cap)."""
    import threading

    sys.path.insert(0, os.path.join(REPO, "tests"))
    import live_util as L
    import pose_util as P

    m = P.reference_metrics()
    if m is None or L.reference_root() is None:
        return {"unavailable": "reference Python not staged (oracle/_ref/py)"}
    ref = L.load_reference()
    import pats_b200.install as inst
    from pats_b200 import pipeline as PL

    ds = P.SyntheticTwoView(n_pairs=n_pairs + 2, hw=(H, W_IMG), n_points=1500, outliers=0.3, load_cost=6)
    out = {"pairs_timed": n_pairs, "unit": "pairs/s"}
    with torch.no_grad():
        real = L.build_model(ref, L.config(if_local=True, merge_new=True, if_outdoor=True), device=dev)
        inst.install(fused=True, attention=True)
        try:
            model = P.PlantedModel(real=real)
            PL.evaluate_pairs_sequential(model, ds, m.compute_pose_error, 1.0, 0.5, device=dev, indices=[n_pairs, n_pairs + 1])  # warm-up
            torch.cuda.synchronize(dev)
            # the GPU stage alone (forward pass incl. host -> device of the images and device -> host of the match lists)
            model.real_matches = 0
            t0 = time.perf_counter()
            for i in range(n_pairs):
                data = PL._collate(ds[i])
                data["image0"], data["image1"] = data["image0"].to(dev), data["image1"].to(dev)
                model(data)["matches_l"].cpu()
            torch.cuda.synchronize(dev)
            out["forward_only_s_per_pair"] = (time.perf_counter() - t0) / n_pairs
            out["forward_matches_per_pair"] = model.real_matches / n_pairs
            box = {}

            def seq():
                t0 = time.perf_counter()
                box["r"] = PL.evaluate_pairs_sequential(model, ds, m.compute_pose_error, 1.0, 0.5, device=dev, indices=range(n_pairs))
                torch.cuda.synchronize(dev)
                box["s"] = time.perf_counter() - t0

            th = threading.Thread(target=seq)  # a fresh thread, like the pipeline's metrics thread: same OpenCV generator state
            th.start()
            th.join()
            stats = {}
            t0 = time.perf_counter()
            par = PL.evaluate_pairs(model, ds, m.compute_pose_error, 1.0, 0.5, device=dev, indices=list(range(n_pairs)), stats=stats)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
        finally:
            inst.uninstall()
    out["sequential"] = {"value": n_pairs / box["s"], "s_per_pair": box["s"] / n_pairs}
    out["pipelined"] = {"value": n_pairs / dt, "s_per_pair": dt / n_pairs, "load_s_per_pair": stats["load_s"] / n_pairs, "pose_s_per_pair": stats["pose_s"] / n_pairs}
    out["speedup"] = box["s"] / dt
    out["pose_errors_identical"] = bool(box["r"] == par)
    return out


def stress_leg(torch, dev, peak):
    """BASELINE.json configs[4] ("stress: 1024x1024 pairs, 4096 coarse patches, 200 Sinkhorn iters, 8 x B200"): the level-1 solve of
    a 1024 x 1024 pair (32 x 32 coarse patches -> one 1025 x 1025 plan, 100 iterations) and the synthetic N = 4096 plan at 200
    iterations, ONE problem per GPU (under `--gpus 8` every rank solves its own).  Both run the grid-cooperative streaming kernel
    (csrc/sinkhorn_grid.cu).  Algorithmic bytes as for `roofline_streaming`: 4 (N+1)^2 (iters + 2).  A single plan of these sizes
    (4 MB / 67 MB) stays in the 126 MB L2 between iterations, so `gbs` may exceed the HBM peak: the bound is L2 / issue, not HBM;
    `frac_of_hbm_peak` is listed because BASELINE.md section 4 defines the figure that way."""
    from pats_b200 import modules as M

    out = {}
    one_d = torch.tensor(1.0, device=dev)
    for name, N, iters, reps in (("pair1024_L1_1025x1025_it100", 1024, 100, 20), ("N4096_it200", 4096, 200, 5)):
        g2 = torch.Generator().manual_seed(SEED + N)
        sc = (0.1 * torch.randn(1, N, N, generator=g2)).to(dev)
        nss = torch.exp((torch.rand(1, 1, N, generator=g2) * 2 - 1) * math.log(16.0)).to(dev)
        for _ in range(3):
            M.log_optimal_transport(sc, one_d, nss, iters)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0_, e1_ in evs:
            e0_.record()
            M.log_optimal_transport(sc, one_d, nss, iters)
            e1_.record()
        torch.cuda.synchronize(dev)
        ms = sorted(a_.elapsed_time(b_) for a_, b_ in evs)[len(evs) // 2]
        nbytes = 4 * (N + 1) * (N + 1) * (iters + 2)
        out[name] = {"ms": ms, "problems_per_s_per_gpu": 1e3 / ms, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "plan_mb": 4 * (N + 1) * (N + 1) / 1e6}
        del sc, nss
    torch.cuda.empty_cache()
    return out


def attention_leg(torch, dev, peak_bf16_tflops):
    """N3: the attention network in front of each matching level (models/modules.py:119-134) at the shapes of ONE 640x480 pair --
    level 1: 1 x 448 x 300, 18 layers; level 2: 300 x 264 x 145, 18 layers; level 3: 4800 x 128 x 65, 10 layers -- the reference's own
    module on this GPU (cuDNN Conv1d = single-pass TF32, cuBLAS FP32 einsum; ~25 ATen launches per layer and side) against
    pats_attentional_gnn_f32 (csrc/gnn.cu) in its default 3xTF32 mode and in single-pass TF32.  Weights: torch's default
    initialisation (seeded); inputs: unit normal.  `useful_tflops` counts each multiply-add of the network once (2 T 9 D^2 per layer
    for the convolutions, 4 n^2 D per problem side for the attention); the tensor peak is half the measured BF16 figure (TF32)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import live_util as L

    if L.reference_root() is None:
        return {"unavailable": "reference Python not staged (oracle/_ref/py)"}
    ref = L.load_reference()
    from pats_b200 import gnn as G

    out = {"tensor_peak_tflops_tf32": peak_bf16_tflops / 2.0, "levels": {}}
    total_ref = total_ours = 0.0
    with torch.no_grad():
        for name, B, D, N, layers in (("L1", 1, 448, 300, 18), ("L2", P2, 264, 145, 18), ("L3", K3, 128, 65, 10)):
            torch.manual_seed(SEED)
            mod = ref.modules.AttentionalGNN(D, ["self", "cross"] * (layers // 2)).eval().to(dev)
            g = torch.Generator().manual_seed(SEED + D)
            x0, x1 = torch.randn(B, D, N, generator=g).to(dev), torch.randn(B, D, N, generator=g).to(dev)

            def timed(fn, reps=3):
                fn()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize(dev)
                return e0.elapsed_time(e1) / reps

            T = 2 * B * N
            flops = layers * (2.0 * T * 9 * D * D + 2 * B * 4.0 * N * N * D)
            rec = {"shape": [B, D, N], "layers": layers, "reference_module_ms": timed(lambda: mod(x0, x1))}
            try:
                for passes in (3, 1):
                    G.set_precision(passes)
                    rec[f"ours_{passes}xtf32_ms"] = timed(lambda: G.attentional_gnn_forward(mod, x0, x1))
            finally:
                G.set_precision(3)
            a0, _ = G.attentional_gnn_forward(mod, x0, x1)
            b0, _ = mod(x0, x1)
            rec["max_abs_diff_vs_reference_module"] = float((a0 - b0).abs().max())
            rec["output_scale"] = float(b0.abs().max())
            rec["useful_tflops"] = flops / (rec["ours_3xtf32_ms"] * 1e-3) / 1e12
            # tensor-core work actually issued in the default mode (three TF32 MMAs per product) over the WHOLE call's time, attention and
            # transpositions included: a lower bound of the GEMM kernels' rate (ncu: tensor pipe 50 % active inside them)
            gemm_flops = layers * 2.0 * T * 9 * D * D
            rec["roofline"] = {"bound": "tensor", "achieved": 3.0 * gemm_flops / (rec["ours_3xtf32_ms"] * 1e-3) / 1e12, "peak": peak_bf16_tflops / 2.0, "unit": "TFLOP/s",
                               "frac": 3.0 * gemm_flops / (rec["ours_3xtf32_ms"] * 1e-3) / 1e12 / (peak_bf16_tflops / 2.0),
                               "note": "TF32 MMA flops issued (3 passes) / time of the whole network call; peak = measured BF16 TFLOP/s / 2"}
            rec["speedup"] = rec["reference_module_ms"] / rec["ours_3xtf32_ms"]
            total_ref += rec["reference_module_ms"]
            total_ours += rec["ours_3xtf32_ms"]
            out["levels"][name] = rec
    out["reference_module_ms_per_pair"] = total_ref
    out["ours_ms_per_pair"] = total_ours
    out["speedup"] = total_ref / total_ours
    return out


def correlation_leg(torch, dev, B):
    """N2: the level-3 correlation in front of the solve (third_layer.py:156-158: einsum('bdn,bdm->bnm') / sqrt(128), then 0.1 *) as the
    reference formulates it (cuBLAS FP32 GEMM + two elementwise passes) against pats_correlation_f32 (tcgen05, 3xTF32), each followed
    by the level-3 Sinkhorn, on B pairs' worth of problems.  Descriptors are synthetic (unit normal)."""
    from pats_b200 import layers as Ly
    from pats_b200 import modules as M

    g = torch.Generator().manual_seed(SEED)
    K = B * K3
    d0 = torch.randn(K, 128, 65, generator=g).to(dev)
    d1 = torch.randn(K, 128, 65, generator=g).to(dev)
    ns = torch.exp((torch.rand(K, 1, 64, generator=g) * 2 - 1) * math.log(2.0)).to(dev)
    one = torch.tensor(1.0, device=dev)
    sc = 0.1 / math.sqrt(128.0)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False

    def timed(fn):
        for _ in range(3):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a_, b_ in evs:
            a_.record()
            fn()
            b_.record()
        torch.cuda.synchronize(dev)
        return sorted(a_.elapsed_time(b_) for a_, b_ in evs)[5]

    try:
        ref_corr = timed(lambda: 0.1 * (torch.einsum('bdn,bdm->bnm', d0, d1) / 128 ** .5))
        our_corr = timed(lambda: Ly.correlation(d0, d1, sc))
        ref_both = timed(lambda: M.log_optimal_transport2(0.1 * (torch.einsum('bdn,bdm->bnm', d0, d1) / 128 ** .5), one, ns, ITERS))
        our_both = timed(lambda: M.log_optimal_transport2(Ly.correlation(d0, d1, sc), one, ns, ITERS))
        Za = M.log_optimal_transport2(Ly.correlation(d0, d1, sc), one, ns, ITERS)
        Zb = M.log_optimal_transport2(0.1 * (torch.einsum('bdn,bdm->bnm', d0, d1) / 128 ** .5), one, ns, ITERS)
        out = {"problems": K, "shape": "[K,128,65] x [K,128,65] -> [K,65,65]", "reference_formulation_ms": ref_corr, "tcgen05_ms": our_corr,
               "reference_formulation_plus_sinkhorn_ms": ref_both, "tcgen05_plus_sinkhorn_ms": our_both,
               "plans_max_abs_diff": float((Za - Zb).abs().max()), "argmax_equal": bool(torch.equal(Za.argmax(2), Zb.argmax(2)) and torch.equal(Za.argmax(1), Zb.argmax(1)))}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    del d0, d1
    torch.cuda.empty_cache()
    return out


def workload_config(pairs_per_step: int) -> dict:
    """The `config` of BOTH arms (key for key: the driver compares them)."""
    return {"workload": WORKLOAD, "scores": "planted (peaked, area-consistent plans; make_inputs)", "pairs_per_step": pairs_per_step, "P2": P2, "K3": K3,
            "sinkhorn_iters": ITERS}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port + oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    import oracle

    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    torch.set_num_threads(cores)
    host = make_inputs(torch, 1, SEED, "planted")  # every step = ONE pair of the step's `pairs_per_step` (they are independent and identical in cost)
    # a full pair costs ~1.3 s on 16 host threads: every step is the whole pair (no sampling, no scaling); a box with few
    # cores falls back to a quarter / an eighth of the level-2 / level-3 problems so that the run stays within minutes
    t0 = time.perf_counter()
    cpu_sample_step(torch, host, 1.0 / 32, 1.0 / 16)  # warm-up (library load, OpenMP pool) and a speed probe
    probe = time.perf_counter() - t0
    frac3, frac2 = (1.0, 1.0) if probe < 1.0 else (1.0 / 8, 1.0 / 4)
    times = []
    desc = ""
    t_wall = time.perf_counter()
    for _ in range(min(args.steps, 20)):
        s, desc, _ = cpu_sample_step(torch, host, frac3, frac2)
        times.append(s)
        if time.perf_counter() - t_wall > 150:  # keep the whole run within a few minutes
            break
    per_pair = sum(times) / len(times)
    value = 1.0 / per_pair
    B = args.pairs_per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1),
        "ms_per_step": per_pair * B * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": f"each timed step = 1 of the step's {B} pairs (independent, equal cost; ms_per_step = {B} x that): " + desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pairs-per-step", type=int, default=DEFAULT_PAIRS_PER_STEP)
    ap.add_argument("--workload", default="planted", choices=["planted", "diffuse"], help="synthetic score model (make_inputs)")
    ap.add_argument("--no-diffuse", action="store_true", help="skip the second, short pass on the diffuse (0.1 * randn) workload")
    ap.add_argument("--no-forward", action="store_true", help="skip the PATS.forward leg (reference on CUDA vs install())")
    ap.add_argument("--no-forward-cpu", action="store_true", help="forward leg: skip the reference PATS.forward on CPU tensors (one pair, ~20 s)")
    ap.add_argument("--streams", type=int, default=1,
                    help="CUDA streams the device-resident steps alternate over (independent pairs overlap; 1 = strictly serial)")
    ap.add_argument("--no-overlap", action="store_true", help="skip the informational two-stream pass")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--wc-staging", action="store_true", help="A/B: write-combined pinned memory for the e2e staging arena")
    ap.add_argument("--no-numa", action="store_true", help="A/B: do not bind the process to the GPU's NUMA node before allocating pinned memory")
    ap.add_argument("--no-chain", action="store_true", help="A/B: plain kernel launches instead of launch chaining (pats_launch_chaining(0))")
    ap.add_argument("--no-handover", action="store_true", help="A/B: no plan hand-over inside the composite calls (pats_plan_handover(0))")
    ap.add_argument("--no-streaming", action="store_true", help="skip the b=32 N=1536 streaming-kernel roofline sample")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the torch-CUDA restatement of the reference's OT")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path is CUDA-only (there is no CPU fallback to time)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned staging buffers of the e2e leg are local
    # to the GPU's PCIe root (the CPU-baseline leg below gets the full CPU set back)
    cpus_before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = {"bound": False, "why": "--no-numa"}
    if not args.no_numa:
        from pats_b200.dist import bind_host_to_gpu

        numa = bind_host_to_gpu(local_rank)
        log(f"[bench] rank {rank}: NUMA binding {numa}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs_per_step

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    gather = None
    if world > 1:
        from pats_b200.dist import gather_match_lists as gather

    if args.no_chain or args.no_handover:
        from pats_b200 import _lib as _l

        if args.no_chain:
            _l.load().pats_launch_chaining(0)
        if args.no_handover:
            _l.load().pats_plan_handover(0)
    kind = args.workload
    steps_by_kind = {}

    def timed_pass(S, n_steps, want_clocks, kind=kind):
        """Times exactly n_steps steps alternating over S CUDA streams (S = 1: strictly serial on the current stream).
        Pairs are independent (evaluate.py:25-35); with S > 1 one batch's single-problem level-1 kernels overlap another
        batch's level-3 work.  Every step is one full pass of the hot path over `pairs_per_step` pairs.  With more than one
        rank the timed region ends with the path's one exchange: EVERY step's match list of every rank is gathered to
        rank 0 (BASELINE config 4: contiguous shards of the pair list, evaluate.py:29-35 computes the metrics in one place)."""
        steps_all = steps_by_kind.setdefault(kind, [])
        while len(steps_all) < S:
            steps_all.append(DeviceStep(torch, dev, B, SEED + rank + 1000 * len(steps_all), kind))
        step = steps_all[0]
        main_stream = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(dev) for _ in range(S)] if S > 1 else [main_stream]
        # inputs + outputs of one step far exceed the 126 MB L2 (level-3 plans alone: 2 x 81 MB per pair) -> no L2 flush needed
        for w in range(max(args.warmup, S)):
            with torch.cuda.stream(streams[w % S]):
                steps_all[w % S].run(streams[w % S].cuda_stream)
        torch.cuda.synchronize(dev)
        nbad = int(step.o["ci_meta"][1].item())
        assert nbad == 0, f"synthetic Compute_imgs inputs produced {nbad} invalid crops"
        kfs = [int(st_.o["gr_total"].item()) for st_ in steps_all[:S]]  # match rows per step (fixed inputs -> fixed count)
        windows = int(step.o["ci_meta"][0].item())
        # this rank's shard of the pair list: n_steps * B pairs; every step appends its match rows [yl, xl, yr, xr] to `acc`
        acc = torch.empty((sum(kfs[i % S] for i in range(n_steps)), 4), dtype=torch.float32, device=dev) if gather is not None else None
        gstats = {}
        def all_lists():
            lists, o2 = [], 0
            for i in range(n_steps):
                lists.append(acc[o2:o2 + kfs[i % S]])
                o2 += kfs[i % S]
            return lists

        if gather is not None:  # warm-up of the exchange with the real sizes: NCCL builds its channels lazily, and the receive buffers come
            gather(all_lists(), max_pairs=n_steps, dst=0)  # out of torch's caching allocator afterwards instead of cudaMalloc
            torch.cuda.synchronize(dev)
        l3_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if (i % L3_SAMPLE_EVERY) == L3_SAMPLE_EVERY - 1 or
                     n_steps < L3_SAMPLE_EVERY else None for i in range(n_steps)]
        sampler = ClockSampler(local_rank)
        if rank == 0 and want_clocks:
            sampler.start()
        e0, e1, eg = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = [torch.cuda.Event() for _ in range(S)]
        barrier()
        t_wall0 = time.perf_counter()
        e0.record(main_stream)
        if S > 1:
            for st in streams:
                st.wait_event(e0)
        launches, off = 0, 0
        for i in range(n_steps):
            with torch.cuda.stream(streams[i % S]):
                st_ = steps_all[i % S]
                launches += st_.run(streams[i % S].cuda_stream, l3_events[i])
                if acc is not None:  # keep this step's list (two device copies; no host sync, no allocation)
                    k = kfs[i % S]
                    acc[off:off + k, :2].copy_(st_.o["ml"][:k])
                    acc[off:off + k, 2:].copy_(st_.o["mr"][:k])
                    off += k
        if S > 1:
            for i, st in enumerate(streams):
                done[i].record(st)
                main_stream.wait_event(done[i])
        eg.record(main_stream)
        if gather is not None:  # one list per step (= per batch of B pairs), all of them, to rank 0
            gather(all_lists(), max_pairs=n_steps, dst=0, stats=gstats)
        e1.record(main_stream)
        barrier()
        t_wall1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1), eg.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop(t_wall0, t_wall1) if (rank == 0 and want_clocks) else None
        l3 = sorted(ev[0].elapsed_time(ev[1]) for ev in l3_events if ev is not None)
        info = {"kf": kfs[0], "windows": windows, "gather_ms": float(ms[1].item()) if gather is not None else 0.0, "gather": gstats,
                "fallbacks": int(step.lib.pats_sinkhorn_fallback_count(1))}
        return float(ms[0].item()), launches, sum(l3) / len(l3), clocks, info

    S = max(1, args.streams)
    from pats_b200 import _lib as _l0

    _l0.load().pats_sinkhorn_fallback_count(1)
    _l0.load().pats_sinkhorn_iterations_skipped(1)
    total_ms, launches, l3_avg_exit, clocks, info = timed_pass(S, args.steps, True)
    skipped = int(_l0.load().pats_sinkhorn_iterations_skipped(1))
    step = steps_by_kind[kind][0]
    kf = info["kf"]
    value = world * B * args.steps / (total_ms * 1e-3)
    # The same pass with the level-3 kernel's fixed-point exit switched off (every problem runs all 100 iterations): this is the
    # launch time the roofline is quoted on (its useful-FMA count assumes every iteration), and the number to compare with round 1.
    _l0.load().pats_sinkhorn_fixed_point_exit(0)
    try:
        n_full = max(3, min(args.steps, 40))
        full_ms, _, l3_avg, _, _ = timed_pass(S, n_full, False)
    finally:
        _l0.load().pats_sinkhorn_fixed_point_exit(1)
    # ... and with the level-3 problems loaded / stored by plain LDG / STG instead of the bulk-copy (TMA engine) staging
    _l0.load().pats_sinkhorn_bulk_staging(0)
    try:
        direct_ms, _, l3_direct, _, _ = timed_pass(S, n_full, False)
    finally:
        _l0.load().pats_sinkhorn_bulk_staging(1)
    bulk = {"enabled": True, "what": "65x65 problems staged by one cp.async.bulk of their 16-byte aligned superset (mbarrier), result formed in place, one bulk "
                                     "store; persistent CTAs; bit-identical (tests/test_gpu_ot.py::test_bulk_staging_is_bit_identical)",
            "l3_ms_per_launch": l3_avg_exit, "l3_ms_per_launch_direct_loads": l3_direct, "value_direct_loads": world * B * n_full / (direct_ms * 1e-3)}
    fp_exit = {"enabled": True, "what": "65x65 kernel leaves its loop at a bitwise fixed point of beta; results identical to the full 100 iterations "
                                        "(tests/test_gpu_ot.py::test_fixed_point_exit_is_bit_identical)",
               "iters_executed_mean": ITERS - skipped / float(B * K3 * (args.steps + max(args.warmup, S))),
               "l3_ms_per_launch": l3_avg_exit, "l3_ms_per_launch_without_exit": l3_avg,
               "value_without_exit": world * B * n_full / (full_ms * 1e-3), "ms_per_step_without_exit": full_ms / n_full}
    overlap = None
    if S == 1 and not args.no_overlap:
        o_ms, _, _, _, _ = timed_pass(2, args.steps, False)
        overlap = {"streams": 2, "value": world * B * args.steps / (o_ms * 1e-3), "unit": "pairs/s", "ms_per_step": o_ms / args.steps,
                   "note": "same steps alternating over two CUDA streams (independent batches overlap); informational"}
    diffuse = None
    if kind == "planted" and not args.no_diffuse and world == 1:
        n_d = max(3, min(args.steps, 20))
        d_ms, _, d_l3, _, d_info = timed_pass(1, n_d, False, "diffuse")
        diffuse = {"workload": WORKLOAD, "scores": "diffuse (0.1 * randn: flat plans, boxes grow to the iteration limit; the round-1 workload)",
                   "value": world * B * n_d / (d_ms * 1e-3), "unit": "pairs/s", "ms_per_step": d_ms / n_d, "steps": n_d, "l3_ms_per_launch": d_l3,
                   "matches_per_pair": d_info["kf"] // B}
        del steps_by_kind["diffuse"]
        torch.cuda.empty_cache()

    # ---- e2e: public API, host buffers ------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        es = E2EStep(torch, dev, B, step.host_inputs, kind, wc=args.wc_staging)
        es.run(2)
        barrier()
        n_e2e = max(3, min(args.steps, 400))  # the same K steps as the device-resident pass (2.3 ms each: the pipeline's fill and drain
        # -- one un-overlapped copy, one un-overlapped compute -- weigh 5 % at 20 steps and 0.5 % at 200)
        t0 = time.perf_counter()
        es.run(n_e2e)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * n_e2e / float(dt.item()), "unit": "pairs/s", "h2d_bytes_per_step": es.h2d_bytes, "d2h_bytes_per_step": es.d2h_bytes,
               "steps": n_e2e}
        # what bounds it: the host->device link.  Time the bare arena copy (same pinned buffer, nothing else running).
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ce0.record()
        for _ in range(5):
            es.d_arena[0].copy_(es.h_arena, non_blocking=True)
        ce1.record()
        torch.cuda.synchronize(dev)
        link_ms = ce0.elapsed_time(ce1) / 5
        e2e["staging"] = es.staging
        e2e["h2d_link_gbs"] = es.h_arena.numel() / (link_ms * 1e-3) / 1e9
        if world > 1:  # every rank's host link, measured while all ranks copy at once: attributes the e2e scaling to the shared host path
            links = torch.zeros(world, device=dev, dtype=torch.float64)
            links[rank] = e2e["h2d_link_gbs"]
            dist.all_reduce(links)
            e2e["h2d_link_gbs_per_rank"] = [round(float(x), 1) for x in links.tolist()]
        e2e["h2d_ms_per_step_alone"] = link_ms
        e2e["note"] = "bound by the host->device copy of the stage inputs (h2d_ms_per_step_alone vs ms_per_step of the device-resident path)"

    fwd_sharded = None
    if world > 1 and not args.no_forward:
        try:
            fwd_sharded = forward_sharded_leg(torch, dist, dev, rank, world)
        except Exception as e:  # noqa: BLE001  (needs the staged reference Python on every rank)
            fwd_sharded = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peak, peak_src = peaks()
        b3 = B * K3
        alg_bytes = b3 * 4 * 65 * 65 * (ITERS + 2)             # streaming model, one pass per iteration + read + write
        compulsory = b3 * 4 * 65 * 65 * 2                      # what an on-chip-resident kernel must move
        achieved = alg_bytes / (l3_avg * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fma_lane_peak = 148 * 128 * sm_mhz * 1e6                # FP32 FMA lanes/s
        fma_done = b3 * 65 * 65 * 2 * (ITERS - 1)               # two FMA passes over the plan per iteration
        flops = 2.0 * fma_done / (l3_avg * 1e-3) / 1e12        # useful FP32 work of the scaling iterations, TFLOP/s
        flops_peak = 2.0 * fma_lane_peak / 1e12                # 148 SMs x 128 FP32 lanes x 2 x f_SM (measured clock of this run)
        roofline = {
            "kernel": "sinkhorn_w65x2_kernel (level-3 OT, two warps per 65x65 problem, %d problems per launch)" % b3,
            # The plan is REGISTER-resident for all 100 iterations: HBM carries only the compulsory read + write (7 % of peak), so
            # the HBM roofline does not bound this kernel.  What does: FP32 issue -- two FFMA passes over the plan per iteration
            # plus the shuffle/FADD butterflies of the row and column sums (ncu: issue slots 51 %, FMA pipe 44 %).  `frac` is the
            # useful-FMA fraction of the FP32 peak; the HBM-bound member of the family is under `roofline_streaming`.
            "bound": "fp32_issue", "achieved": flops, "peak": flops_peak, "unit": "TFLOP/s", "frac": flops / flops_peak,
            "traffic": ncu_traffic(), "peak_source": "148 SM x 128 FP32 lanes x 2 x SM clock sampled during this run (%.0f MHz)" % sm_mhz,
            "ms_per_launch": l3_avg, "useful_fma": fma_done,
            "hbm": {"compulsory_bytes": compulsory, "frac_of_hbm_peak": compulsory / (l3_avg * 1e-3) / 1e9 / peak, "hbm_peak_gbs": peak, "peak_source": peak_src,
                    "streaming_model_bytes": alg_bytes, "streaming_model_gbs": achieved,
                    "note": "streaming model (SURVEY.md 8d: one HBM pass per iteration) divided by the launch time exceeds the HBM peak "
                            "because the kernel does not stream; it is listed for reference only"},
            "share_of_step": l3_avg * n_full / full_ms,
            "timed": "with the fixed-point exit OFF (all 100 iterations of every problem); see `fixed_point_exit` for the default",
        }
        # ---- the streaming member of the kernel family against the HBM roofline it is really bound by ------------------
        # BASELINE.json configs[2]: b = 32, N = 1536 (-> 1537 x 1537 plans, 302 MB in + 302 MB out: nothing fits on chip,
        # the plan is read from HBM once per iteration).  Algorithmic bytes per SURVEY.md 8(d): b*4*(N+1)^2*(iters+2).
        streaming = None
        if world == 1 and not args.no_streaming:
            from pats_b200 import modules as M

            g2 = torch.Generator().manual_seed(SEED)
            bs, Ns = 32, 1536
            sc = (0.1 * torch.randn(bs, Ns, Ns, generator=g2)).to(dev)
            nss = torch.exp((torch.rand(bs, 1, Ns, generator=g2) * 2 - 1) * math.log(16.0)).to(dev)
            one_d = torch.tensor(1.0, device=dev)
            for _ in range(3):
                M.log_optimal_transport(sc, one_d, nss, ITERS)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for e0_, e1_ in evs:
                e0_.record()
                M.log_optimal_transport(sc, one_d, nss, ITERS)
                e1_.record()
            torch.cuda.synchronize(dev)
            ms_s = sum(a_.elapsed_time(b_) for a_, b_ in evs) / len(evs)
            bytes_s = bs * 4 * (Ns + 1) * (Ns + 1) * (ITERS + 2)
            streaming = {"kernel": "sinkhorn_grid_kernel (plans beyond 512 x 512: rows split over co-resident CTAs, one HBM pass per iteration)",
                         "workload": "BASELINE.json configs[2]: b=32, N=1536, 100 it", "bound": "hbm", "achieved": bytes_s / (ms_s * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": bytes_s / (ms_s * 1e-3) / 1e9 / peak, "ms_per_launch": ms_s,
                         "algorithmic_bytes": bytes_s, "traffic": ncu_traffic("grid_dram_bytes_per_launch")}
            del sc, nss
            torch.cuda.empty_cache()
        stress = None
        if not args.no_streaming:
            stress = stress_leg(torch, dev, peak)
        corr = None
        if world == 1 and not args.no_streaming:
            try:
                corr = correlation_leg(torch, dev, B)
            except Exception as e:  # noqa: BLE001
                corr = {"unavailable": f"{type(e).__name__}: {e}"}
        # ---- the reference's own formulation (log-domain, ~6 ATen ops per iteration: modules.py:137-182) on this GPU -------
        torch_cuda = None
        if world == 1 and not args.no_torch_baseline:
            def torch_ot2(sx_, ns_, iters):
                b_, m_, n_ = sx_.shape
                nsum = ns_.sum(2).reshape(b_)
                norm = -(float(m_ - 1) + nsum).log()
                lnu = torch.cat([ns_.reshape(b_, -1).log() + norm[:, None], (math.log(float(m_ - 1)) + norm)[:, None]], 1)
                lmu = torch.cat([norm[:, None].expand(b_, m_ - 1), (nsum.log() + norm)[:, None]], 1)
                u, v = torch.zeros_like(lmu), torch.zeros_like(lnu)
                for _ in range(iters):
                    u = lmu - torch.logsumexp(sx_ + v[:, None, :], 2)
                    v = lnu - torch.logsumexp(sx_ + u[:, :, None], 1)
                return sx_ + u[:, :, None] + v[:, None, :] - norm[:, None, None]

            tc = {}
            for name, sk, nk in (("ot2_300x145", "l2_scores", "l2_ns"), ("ot3_4800x65", "l3_scores", "l3_ns")):
                torch_ot2(step.i[sk], step.i[nk], 3)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                torch_ot2(step.i[sk], step.i[nk], ITERS)
                eb.record()
                torch.cuda.synchronize(dev)
                tc[name + "_ms"] = ea.elapsed_time(eb)
            torch_cuda = {"what": "log_optimal_transport2 restated with the reference's ATen ops (logsumexp per half-iteration), same GPU, same inputs, "
                                  "CUDA events; informational -- the contract's baseline is the CPU arm", **tc,
                          "ours_ot3_ms": l3_avg, "speedup_ot3": tc["ot3_4800x65_ms"] / l3_avg}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            import oracle

            if cpus_before is not None and numa.get("bound"):
                try:
                    os.sched_setaffinity(0, cpus_before)  # the CPU arm uses every host core
                except OSError:
                    pass

            cores = os.cpu_count() or 1
            oracle.set_num_threads(cores)
            torch.set_num_threads(cores)
            hin = make_inputs(torch, 1, SEED, kind)
            t0 = time.perf_counter()
            cpu_sample_step(torch, hin, 1.0 / 32, 1.0 / 16)  # warm-up and speed probe
            full = (time.perf_counter() - t0) < 1.0
            runs, t_wall = [], time.perf_counter()
            while len(runs) < 8 and (not runs or time.perf_counter() - t_wall < 12.0):  # about 10 s of CPU work
                runs.append(cpu_sample_step(torch, hin, 1.0 if full else 1.0 / 4, 1.0 if full else 1.0 / 2))
            secs = sum(r[0] for r in runs) / len(runs)
            desc, parts = runs[-1][1], runs[-1][2]
            cpu = {"value": 1.0 / secs, "unit": "pairs/s", "cores": oracle.num_threads(), "kind": "port",
                   "sample": f"mean of {len(runs)} runs of: {desc}", "seconds_per_pair": secs, "parts_s": {k: round(v, 4) for k, v in parts.items()}}
        fwd = None
        if world == 1 and not args.no_forward:
            try:
                fwd = forward_leg(torch, dev, args)
            except Exception as e:  # noqa: BLE001  (the leg needs the staged reference Python, oracle/_ref/py)
                fwd = {"unavailable": f"{type(e).__name__}: {e}"}
        fwd_global = None
        if world == 1 and not args.no_forward:
            try:
                fwd_global = forward_leg(torch, dev, args, if_local=False, arms=("reference_cuda", "installed_fused", "installed_fused_attention"))
            except Exception as e:  # noqa: BLE001
                fwd_global = {"unavailable": f"{type(e).__name__}: {e}"}
        att = None
        if world == 1 and not args.no_forward:
            try:
                att = attention_leg(torch, dev, tensor_peak_bf16())
            except Exception as e:  # noqa: BLE001
                att = {"unavailable": f"{type(e).__name__}: {e}"}
        tail = None
        if world == 1 and not args.no_forward:
            try:
                tail = tail_leg(torch, dev)
            except Exception as e:  # noqa: BLE001
                tail = {"unavailable": f"{type(e).__name__}: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B),
            "run": {"streams": S, "host_numa": numa, "l2_flush": "not needed: one step streams > 300 MB of distinct plans per pair (L2 = 126 MB)",
                    "matches_per_pair": kf // B, "windows_per_pair": info["windows"] // B, "sinkhorn_fallbacks": info["fallbacks"],
                    "pairs_per_rank": B * args.steps,
                    "exchange": "after the pair loop every step's match list of every rank is gathered to rank 0 (2 collectives, 1 host sync); inside the timed region",
                    "gather_ms": info["gather_ms"], "gather": info["gather"]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_streaming": streaming, "stress": stress, "correlation": corr, "cpu_baseline": cpu, "torch_cuda": torch_cuda,
            "fixed_point_exit": fp_exit, "bulk_staging": bulk, "overlap": overlap, "diffuse": diffuse, "forward": fwd, "forward_global": fwd_global, "attention": att, "forward_sharded": fwd_sharded, "tail": tail,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
