"""Shared by tests/test_oracle_trace.py (CPU) and tests/test_gpu_trace.py (GPU): reading tests/golden/trace_*.npz
(call records of the unmodified reference `PATS.forward`, tests/golden/make_trace.py) and the comparison rules.

Tolerances (the same ones tests/test_oracle_golden.py uses against the per-function fixtures):
    integers, booleans, byte copies, match lists ............ bit-exact
    transport plans ......................................... |d| <= 1e-4 (north_star) + 2e-6 |ref|: the random-init network of the
                                                              trace produces scores up to 4e7, where the spacing of f32 itself
                                                              (7.8e-3 at 1e5) exceeds 1e-4; entries of ordinary magnitude are held
                                                              to 1e-4 absolute, the huge ones to ~16 ulp
    crop + bilinear resize (values 0..255) .................. |d| <= 6e-5 (ATen's vectorised CPU lerp differs in the last ulp)
    area expansion: average_point / scales .................. rtol 2e-5 (+ atol 2e-5)
                    whole_cost (trust score) ................ rtol 2e-4, atol 2e-6  (difference of long f32 sums)
                    core_cost ............................... rtol 2e-3, atol 2e-5  (cancellation)
    Compute_result: mkpts0_f exact, mkpts1_f ................ rtol 1e-5, atol 2e-5
"""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TAGS = ("global", "local", "mergeold", "portrait", "big", "cond", "condlocal")   # replayed on the GPU (tests/test_gpu_trace.py) and on the oracle

EXACT = "exact"
OT = ("ot", 1e-4, 2e-6)
_EXPAND = [(2e-4, 2e-6), (2e-3, 2e-5), (2e-5, 2e-5), (2e-5, 1e-6), (2e-5, 1e-6), EXACT]
_EST = [(2e-4, 2e-6), (2e-5, 2e-5), (2e-5, 1e-6), (2e-5, 1e-6), EXACT, EXACT]
RULES = {
    "log_sinkhorn_iterations": OT,
    "log_optimal_transport": OT,
    "log_optimal_transport2": OT,
    "tensor_resize": (0.0, 6e-5),
    "origin_extract": EXACT,
    "Compute_imgs": [EXACT, (0.0, 6e-5), EXACT, EXACT, EXACT],
    "Iterative_expand_matrix": _EXPAND,
    "FirstLayer.est_position": _EST,
    "SecondLayer.est_position": _EST,
    "SecondLayer.merge_patches_new": EXACT,
    "SecondLayer.merge_patches_old": EXACT,
    "ThirdLayer.Compute_result": [EXACT, (1e-5, 2e-5)],
    "get_result": EXACT,
    "split_patches": EXACT,
}


def path_of(tag):
    return os.path.join(HERE, "golden", f"trace_{tag}.npz")


def load(tag):
    z = np.load(path_of(tag))
    meta = json.loads(bytes(z["__schema__"]).decode())
    return z, meta


def records(tags=TAGS):
    out = []
    for tag in tags:
        if os.path.exists(path_of(tag)):
            _, meta = load(tag)
            out += [(tag, c["seq"], c["name"]) for c in meta["calls"]]
    return out


def ids(recs):
    return [f"{t}-{s:02d}-{n}" for t, s, n in recs]


def decode(node, z, dev=None, as_numpy=False):
    """Schema node -> python value; tensors become torch tensors on `dev` (or numpy arrays with as_numpy)."""
    t = node["t"]
    if t == "tensor":
        a = np.ascontiguousarray(z[node["key"]])
        if as_numpy:
            if node.get("packed"):
                a = a.astype(node["dtype"])
            return a.reshape(node["shape"])
        import torch

        x = torch.from_numpy(a)
        if node.get("packed"):
            x = x.to(getattr(torch, node["dtype"]))
        x = x.reshape(node["shape"])
        return x.to(dev) if dev is not None else x
    if t in ("list", "tuple"):
        items = [decode(v, z, dev, as_numpy) for v in node["items"]]
        return items if t == "list" else tuple(items)
    if t == "dict":
        return {k: decode(v, z, dev, as_numpy) for k, v in node["items"].items()}
    if t in ("none", "self"):
        return None
    if t == "device":
        if as_numpy:
            return None
        import torch

        return torch.device(dev if dev is not None else "cpu")
    return node["v"]


def _np(x):
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def compare_leaf(label, got, want, rule):
    g, w = _np(got), _np(want)
    assert tuple(g.shape) == tuple(w.shape), f"{label}: shape {tuple(g.shape)} != {tuple(w.shape)}"
    if rule == EXACT or w.dtype.kind in "biu":
        if g.dtype != w.dtype:
            g = g.astype(w.dtype)
        nd = int((g != w).sum()) if not (w.dtype.kind == "f") else int(((g != w) & ~(np.isnan(g) & np.isnan(w))).sum())
        assert nd == 0, f"{label}: {nd} of {w.size} entries differ (bit-exact expected)"
        return 0.0
    gf, wf = g.astype(np.float64), w.astype(np.float64)
    fin = np.isfinite(wf)
    assert np.array_equal(fin, np.isfinite(gf)), f"{label}: non-finite pattern differs"
    assert np.array_equal(gf[~fin], wf[~fin], equal_nan=True), f"{label}: inf/nan values differ"
    d = np.abs(gf[fin] - wf[fin])
    if rule[0] == "ot":
        excess = d - (rule[1] + rule[2] * np.abs(wf[fin]))
        mx = float(d.max()) if d.size else 0.0
        assert not (excess > 0).any(), f"{label}: {int((excess > 0).sum())} entries beyond {rule[1]:g} + {rule[2]:g}|ref| (max |d| = {mx:.3e})"
        return mx
    rtol, atol = rule[0], rule[1]
    bad = d > atol + rtol * np.abs(wf[fin])
    if len(rule) > 2 and bad.any():  # (rtol, atol, budget): a float output that sits behind a discrete decision (a box that grew one step
        # further, a neighbourhood centred one cell over) may differ outright in at most `budget` of its entries
        assert bad.mean() <= rule[2], f"{label}: {int(bad.sum())} of {d.size} entries differ ({bad.mean():.2e} > budget {rule[2]:g})"
        return float(d[~bad].max()) if (~bad).any() else 0.0
    assert not bad.any(), f"{label}: {int(bad.sum())} of {d.size} entries beyond rtol {rtol:g} / atol {atol:g} (max |d| {float(d.max()):.3e})"
    return float(d.max()) if d.size else 0.0


def compare(name, got, want, rule=None, label=None):
    """Walk the reference's return structure; `rule` is EXACT, OT, (rtol, atol) or a per-output list of those."""
    rule = RULES[name] if rule is None else rule
    label = label or name
    if isinstance(want, (list, tuple)):
        assert isinstance(got, (list, tuple)) and len(got) >= len(want), f"{label}: structure differs"
        for i, w in enumerate(want):
            r = rule[i] if isinstance(rule, list) else rule
            compare(name, got[i], w, r, f"{label}[{i}]")
    elif want is None:
        assert got is None, label
    elif isinstance(want, (bool, int, float, str)):
        gv = got.item() if hasattr(got, "item") else got
        assert gv == want, f"{label}: {gv} != {want}"
    else:
        compare_leaf(label, got, want, rule)


def ot_reference_torch(name, scores, alpha, ns, iters):
    """models/modules.py:137-182 restated with the same ATen ops (cat / logsumexp / broadcast adds, u before v), in the dtype
    and on the device of `scores`.  Test infrastructure: executed on the GPU it is "the reference run on CUDA tensors", the
    oracle of record of SURVEY.md section 8c, for the ill-conditioned records below; in f64 it measures the reference's own
    f32 rounding error (tools/triage_trace.py)."""
    import torch

    b, m, n = scores.shape
    if name == "log_optimal_transport":
        ms = scores.new_tensor(float(m))
        Z = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], -1), torch.cat([alpha.expand(b, 1, n), alpha.expand(b, 1, 1)], -1)], 1)
        norm = -(ms + ns.sum(dim=2)).log()
        log_nu = torch.cat([ns.log()[:, 0] + norm, ms.log().reshape(1, 1).expand(b, 1) + norm], dim=1)
        log_mu = torch.cat([norm.expand(b, m), ns.sum(dim=2).log() + norm], dim=1)
    else:
        Z = scores
        ms = ((m - 1) * alpha).to(scores)
        norm = -(ms + ns.sum(dim=2)).log()
        log_nu = torch.cat([ns.log()[:, 0] + norm, ms.log().reshape(1, 1).expand(b, 1) + norm], dim=1)
        log_mu = torch.cat([norm.expand(b, m - 1), ns.sum(dim=2).log() + norm], dim=1)
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1) - norm.reshape(b, 1, 1)


ILL_CONDITIONED_SPACING = 1e-5  # f32 spacing at max |score|: beyond it the potentials u, v are rounded more coarsely than the 1e-4 bar


def compare_plan_on_gpu(name, args, got, want):
    """Plan comparison of the GPU replay.  Well-conditioned records (every record a trained or conditioned network produces):
    |d| <= 1e-4 + 2e-6 |ref| against the stored reference output, no exceptions.

    Ill-conditioned records (random-init network, |scores| 1e6 .. 4e7; the f32 spacing there is 0.06 .. 4): the iteration is a
    chain of f32 roundings of numbers ~1e6 whose differences are the result, so the stored values (reference on CPU tensors)
    and the reference on CUDA tensors -- the oracle of record, SURVEY.md section 8c -- can differ by a rounding step wherever
    ATen's two logsumexp kernels sum in a different order (measured: trace `portrait`, record 6: one entry of 63 075 differs
    by 0.0078; the same code in f64 differs from either by 11.3; profiles/r02_triage_trace.json).  For these records the
    replacement must (a) agree with the reference's formulation executed in f32 ON THIS GPU within the same bound everywhere,
    and (b) deviate from the stored CPU values only at entries where that CUDA execution of the reference deviates too."""
    import torch

    scores = args[0]
    g, w = got.double(), want.to(got.device).double()
    lim = OT[1] + OT[2] * w.abs()
    d = (g - w).abs()
    fin = torch.isfinite(w)
    assert torch.equal(fin, torch.isfinite(g)), f"{name}: non-finite pattern differs"
    beyond = (d > lim) & fin
    if not bool(beyond.any()):
        return "stored"
    smax = float(scores.abs().max())
    spacing = float(np.spacing(np.float32(smax)))
    assert spacing > ILL_CONDITIONED_SPACING, (
        f"{name}: {int(beyond.sum())} entries beyond {OT[1]:g} + {OT[2]:g}|ref| on a well-conditioned record (max |score| {smax:.3g}, max |d| {float(d[fin].max()):.3e})")
    iters = int(args[3]) if len(args) > 3 else 100
    rc = ot_reference_torch(name, scores.float(), args[1].float().to(scores.device), args[2].float(), iters).double()
    d_cuda = (g - rc).abs()
    lim_c = OT[1] + OT[2] * rc.abs()
    assert not bool(((d_cuda > lim_c) & fin).any()), (
        f"{name}: ill-conditioned record (f32 spacing {spacing:g}): {int(((d_cuda > lim_c) & fin).sum())} entries differ from the reference formulation "
        f"executed on this GPU (max |d| {float(d_cuda[fin].max()):.3e})")
    ref_dev = ((rc - w).abs() > 0.5 * lim) & fin
    assert not bool((beyond & ~ref_dev).any()), (
        f"{name}: {int((beyond & ~ref_dev).sum())} entries differ from the stored CPU reference where the reference on CUDA agrees with it")
    return "cuda-reference"


def get_path(root, path):
    x = root
    for p in path:
        x = x[p]
    return x


def check_argmax_parity(name, got, want):
    """Row / column argmax of a plan must agree wherever the reference's own top-2 gap exceeds the plan tolerance."""
    g, w = _np(got), _np(want)
    for axis in (1, 2):
        gi, wi = g.argmax(axis), w.argmax(axis)
        diff = gi != wi
        if diff.any():
            top2 = np.sort(w, axis=axis)
            top2 = np.take(top2, [-1, -2], axis=axis)
            gap = np.abs(np.take(top2, 0, axis=axis) - np.take(top2, 1, axis=axis))
            assert (gap[diff] <= 2e-4).all(), f"{name}: argmax over axis {axis} differs where the reference's gap exceeds the tolerance"
