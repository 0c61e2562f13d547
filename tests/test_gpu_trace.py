"""Drop-in replay: the hot-path calls of the UNMODIFIED reference `PATS.forward`, through pats_b200's replacements.

tests/golden/trace_{global,local}.npz (tests/golden/make_trace.py, generated in the container that has /root/reference)
hold, for every hot-path name that `pats_b200.install` rebinds, the arguments exactly as models/pats.py, first_layer.py,
second_layer.py and third_layer.py pass them during a forward pass over the 640x480 synthetic pair (if_local False and
True), the reference's return value, and the post-call value of arguments the reference mutates in place.  Each record
is replayed on the GPU through the object `install()` binds at that site -- what the unmodified caller would reach --
and compared under the rules of tests/trace_util.py (integers / booleans / match lists bit-exact, plans within 1e-4).
"""
from __future__ import annotations

import pytest

import trace_util as T

pytestmark = pytest.mark.gpu
ALL = T.records()


def _replacement(name):
    import pats_b200.install as inst

    if "." in name:
        cls, meth = name.split(".")
        for (_, c, f), repl in inst._METHODS.items():
            if c == cls and f == meth:
                return repl
        raise KeyError(name)
    if name == "tensor_resize":  # bound as a module with one function (utils/utils.py:17, :1385)
        return inst._TABLE["utils.utils"]["tensor_resize"].tensor_resize
    for table in inst._TABLE.values():
        if name in table:
            return table[name]
    raise KeyError(name)


def test_trace_fixtures_cover_the_install_table():
    import pats_b200.install as inst

    assert ALL, "tests/golden/trace_*.npz missing (python tests/golden/make_trace.py)"
    names = {n for _, _, n in ALL}
    bound = {n for table in inst._TABLE.values() for n in table} | {f"{c}.{m}" for (_, c, m) in inst._METHODS}
    # merge_patches_old is the merge_new=False branch (never taken with the reference's configs/*.yaml)
    missing = bound - names - {"SecondLayer.merge_patches_old"}
    assert not missing, f"no trace record for {sorted(missing)}"


@pytest.mark.parametrize("tag,seq,name", ALL, ids=T.ids(ALL))
def test_replay_reference_call(tag, seq, name, monkeypatch):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    z, meta = T.load(tag)
    c = meta["calls"][seq]
    assert c["name"] == name
    args = T.decode(c["args"], z, dev)
    kwargs = T.decode(c["kwargs"], z, dev)
    want = T.decode(c["out"], z, None)
    if "merge_patches" in name:  # the records hold the reference on CPU tensors: first-minimum ties (pats_b200.layers.MERGE_TIE_BREAK)
        import pats_b200.layers as Ly

        monkeypatch.setattr(Ly, "MERGE_TIE_BREAK", "first")
    got = _replacement(name)(*args, **kwargs)
    torch.cuda.synchronize()
    if name == "split_patches":  # (cycle_num, [[lo,hi]...], [[head,tail]...]) of python ints / 0-dim tensors
        norm = lambda r: [int(r[0]), [[int(v) for v in row] for row in r[1]], [[int(v) for v in row] for row in r[2]]]  # noqa: E731
        assert norm(got) == norm(want)
        return
    if name == "ThirdLayer.Compute_result":  # whole_loss (3rd value) is discarded by the only caller, third_layer.py:160
        want = tuple(want[:2])
    if name in ("log_optimal_transport", "log_optimal_transport2"):
        T.compare_plan_on_gpu(name, args, got, want)
    elif name == "Compute_imgs":  # x / y_scale_new: the stored CPU run divides by 96, CUDA ATen multiplies by the reciprocal (1 ulp)
        T.compare(name, got, want, [T.EXACT, (0.0, 6e-5), (1.2e-7, 0.0), (1.2e-7, 0.0), T.EXACT])
    else:
        T.compare(name, got, want)
    # arguments the reference mutates in place (second_layer.py:194-207: trust_score, if_nomatching1_L2, scores_back)
    for m in c["mutated"]:
        path = m["path"]
        if name.endswith("merge_patches_old") and path[:2] == [0, 6]:
            continue  # the reference's final content of this argument is unobservable (pats.py:37 keeps the return value); INTEGRATION.md section 5
        root = args if path[0] == 0 else kwargs
        T.compare(name, T.get_path(root, path[1:]), T.decode(m["value"], z, None), T.EXACT, f"{name}<arg {path[1:]} after the call>")
    if name in ("log_optimal_transport", "log_optimal_transport2"):
        T.check_argmax_parity(name, got, want)
