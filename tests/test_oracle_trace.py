"""Pins the CPU oracle on the call records of the unmodified reference `PATS.forward` (tests/golden/trace_*.npz,
tests/golden/make_trace.py): the same inputs the reference's own callers produced, the reference's own results.  CPU only.
Comparison rules: tests/trace_util.py."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
import trace_util as T

ALL = T.records(T.TAGS)


def _kw(kwargs, args, pos, name, default):
    if name in kwargs:
        return kwargs[name]
    return args[pos] if len(args) > pos else default


def _expand(args, kwargs):
    scores_in, sx, sy, limitation, ranges, positions = args[:6]
    # the TRUE grid: limitation = [0, height, 0, width] (first_layer.py:165, second_layer.py:246).  The reference itself
    # rebuilds (height, width) from ranges / positions (utils.py:1181), which swaps them for portrait grids -- a quirk the
    # oracle reproduces internally from the true grid.
    gh, gw = int(limitation[1]), int(limitation[3])
    assert gh * gw == positions.shape[0] and ranges.shape[0] == max(gh, gw)
    lb = _kw(kwargs, args, 6, "lower_bound", 1e-3)
    it = _kw(kwargs, args, 8, "iter_num", 15)
    whole, core, avg, xs, ys, bound, _ = oracle.iterative_expand_matrix(scores_in, sx, sy, gh, gw, lower_bound=lb, iter_num=it)
    return whole, core, avg, xs, ys, bound


def _est(args, level):
    if level == 1:
        _, scores, scale, image_shape, patch_scale = args
        sx = sy = scale
        it, lb = 15, 1e-5
    else:
        _, scores, sx, sy, image_shape, patch_scale = args
        it, lb = 8, 1e-3
    gh, gw = image_shape[0] // patch_scale, image_shape[1] // patch_scale
    nm1, nm2 = oracle.est_nomatching(scores, gh * gw)
    with np.errstate(over="ignore"):
        whole, _, avg, xs, ys, _, _ = oracle.iterative_expand_matrix(np.exp(scores.astype(np.float32)), sx, sy, gh, gw, lower_bound=lb, iter_num=it)
    return whole, avg, xs, ys, nm1, nm2


def run_oracle(name, args, kwargs):
    if name == "log_sinkhorn_iterations":
        return oracle.log_sinkhorn_iterations(args[0], args[1], args[2], _kw(kwargs, args, 3, "iters", 100))
    if name == "log_optimal_transport":
        return oracle.log_optimal_transport(args[0], float(args[1]), args[2], _kw(kwargs, args, 3, "iters", 100))
    if name == "log_optimal_transport2":
        return oracle.log_optimal_transport2(args[0], float(args[1]), args[2], _kw(kwargs, args, 3, "iters", 100))
    if name == "tensor_resize":
        return oracle.tensor_resize(args[0].astype(np.float32), args[1])
    if name == "origin_extract":
        return oracle.origin_extract(args[0], args[1], args[2], args[3])
    if name == "Compute_imgs":
        nl, nr, xs, ys, avg, _ = oracle.compute_imgs(*args[:6], width=kwargs.get("width", 20), height=kwargs.get("height", 15))
        return nl, nr, xs, ys, avg
    if name == "split_patches":
        return oracle.split_patches(args[0], args[1], args[2], _kw(kwargs, args, 3, "max_once_used", 350))
    if name == "Iterative_expand_matrix":
        return _expand(args, kwargs)
    if name == "FirstLayer.est_position":
        return _est(args, 1)
    if name == "SecondLayer.est_position":
        return _est(args, 2)
    if name.startswith("SecondLayer.merge_patches"):
        _, patch_num, trust, shape, nm1, nm2, sb = args
        out, sb_out, trust_after, nm2_after = oracle.merge_patches(name.endswith("new"), trust, [int(shape[0]), int(shape[1])], nm1, nm2, sb)
        return (out, sb_out), {2: trust_after, 5: nm2_after, 6: sb_out}
    if name == "ThirdLayer.Compute_result":
        _, scores, W, Tt, sx, sy, p_s, p_t = args[:8]
        m0, m1, _ = oracle.third_compute_result(scores, sx, sy, p_s, p_t)
        return m0, m1
    if name == "get_result":
        batch, nm, pts, scale, patch_size = args[:5]
        return oracle.get_result(nm, pts, scale, patch_size)
    raise KeyError(name)


def test_traces_present():
    assert ALL, "tests/golden/trace_*.npz missing (python tests/golden/make_trace.py)"


@pytest.mark.parametrize("tag,seq,name", ALL, ids=T.ids(ALL))
def test_oracle_matches_reference_call(tag, seq, name):
    z, meta = T.load(tag)
    c = meta["calls"][seq]
    args = T.decode(c["args"], z, as_numpy=True)
    kwargs = T.decode(c["kwargs"], z, as_numpy=True)
    want = T.decode(c["out"], z, as_numpy=True)
    got = run_oracle(name, args, kwargs)
    after = None
    if isinstance(got, tuple) and len(got) == 2 and isinstance(got[1], dict):
        got, after = got
    if name == "split_patches":
        norm = lambda r: [int(r[0]), [[int(v) for v in row] for row in r[1]], [[int(v) for v in row] for row in r[2]]]  # noqa: E731
        assert norm(got) == norm(want)
        return
    if name == "ThirdLayer.Compute_result":
        want = tuple(want[:2])
    T.compare(name, got, want)
    for m in c["mutated"]:
        path = m["path"]
        if name.endswith("merge_patches_old") and path[1] == 6:
            # second_layer.py:157 writes this chunk's scores into the caller's scores_back and then returns a NEW zero tensor
            # (:186); the only caller rebinds the name to the return value (pats.py:37), so the argument's final content is
            # unobservable.  The replacement zeroes the argument and returns it -- the return value is what is compared above.
            continue
        assert path[0] == 0 and after is not None and path[1] in after, f"{name}: the reference mutated argument {path} in place"
        T.compare(name, after[path[1]], T.decode(m["value"], z, as_numpy=True), T.EXACT, f"{name}<arg {path[1:]} after the call>")
    if name in ("log_optimal_transport", "log_optimal_transport2"):
        T.check_argmax_parity(name, got, want)
