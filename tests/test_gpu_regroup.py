"""GPU parity tests: area expansion (a8/a9), window regrouping (a11), match assembly (a14), third-layer result (a13).
Integer / boolean / ordering results are bit-exact against the committed reference outputs and the CPU oracle."""
import math

import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _ranges_positions(h, w, dev):
    # what Compute_positions_and_ranges (utils/utils.py:1527) hands to Iterative_expand_matrix; only their shapes are read
    s = max(h, w)
    return torch.zeros(s, s, device=dev), torch.zeros(h * w, 2, device=dev)


@pytest.mark.parametrize("tag,gh,gw,lb,it", [("L1", 15, 20, 1e-5, 15), ("L2", 12, 12, 1e-3, 8)])
def test_iterative_expand_matrix_golden(dev, tag, gh, gw, lb, it):
    from pats_b200 import utils as U

    g = load_golden("expand")
    Z = T(g[tag + "_Z"], dev)
    rng, pos = _ranges_positions(gh, gw, dev)
    lim = torch.tensor([0, gh, 0, gw], device=dev)
    whole, core, avg, xs, ys, bound = U.Iterative_expand_matrix(Z.exp(), T(g[tag + "_scalex"], dev), T(g[tag + "_scaley"], dev), lim, rng, pos,
                                                                height=gh, width=gw, iter_num=it, lower_bound=lb)
    assert bound.dtype == torch.int64
    assert np.array_equal(bound.cpu().numpy(), g[tag + "_bound"])
    np.testing.assert_allclose(avg.cpu().numpy(), g[tag + "_average_point"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(xs.cpu().numpy(), g[tag + "_x_scale"], rtol=2e-5)
    np.testing.assert_allclose(ys.cpu().numpy(), g[tag + "_y_scale"], rtol=2e-5)
    np.testing.assert_allclose(whole.cpu().numpy(), g[tag + "_whole_cost"].reshape(whole.shape), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(core.cpu().numpy(), g[tag + "_core_cost"].reshape(core.shape), rtol=2e-3, atol=2e-5)
    nm1, nm2 = U.est_nomatching(Z, gh * gw)
    assert np.array_equal(nm1.cpu().numpy(), g[tag + "_nm1"]) and np.array_equal(nm2.cpu().numpy(), g[tag + "_nm2"])


def _planted(g, b, gh, gw, sharp, noise):
    n = gh * gw
    ys, xs = torch.meshgrid(torch.arange(gh).float(), torch.arange(gw).float(), indexing="ij")
    src = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
    out = []
    for _ in range(b):
        A = torch.eye(2) * (0.6 + 0.8 * torch.rand(1, generator=g)) + 0.1 * torch.randn(2, 2, generator=g)
        t = torch.randn(2, generator=g) * 1.5
        ctr = torch.tensor([gh / 2.0, gw / 2.0])
        warped = (src - ctr) @ A.T + ctr + t
        d2 = ((warped[:, None, :] - src[None, :, :]) ** 2).sum(-1)
        out.append(-d2 / sharp + noise * torch.randn(n, n, generator=g))
    return torch.stack(out)


@pytest.mark.parametrize("b,gh,gw,lb,it", [(300, 12, 12, 1e-3, 8), (2, 15, 20, 1e-5, 15), (1, 20, 15, 1e-5, 15), (3, 32, 32, 1e-5, 15)])
def test_iterative_expand_matrix_vs_oracle(dev, b, gh, gw, lb, it):
    """Full level-2 batch (P=300 windows of 12x12), level-1 plans, a portrait grid (the reference's width/height quirk)
    and the 1024x1024-pair grid (32x32): boxes bit-exact against the oracle, f32 outputs equal to the last ulps."""
    from pats_b200 import modules as M
    from pats_b200 import utils as U

    g = torch.Generator().manual_seed(50 + gh)
    n = gh * gw
    s = 0.5 * _planted(g, b, gh, gw, 2.0, 0.5)
    s = torch.cat([torch.cat([s, torch.zeros(b, n, 1)], 2), torch.zeros(b, 1, n + 1)], 1)
    sx = torch.exp((torch.rand(b, n, 1, generator=g) * 2 - 1) * 1.1)
    sy = torch.exp((torch.rand(b, n, 1, generator=g) * 2 - 1) * 1.1)
    Z = M.log_optimal_transport2(s.to(dev), 1.0, (sx * sy).reshape(b, 1, n).to(dev), 100)
    scores = Z.exp()
    rng, pos = _ranges_positions(gh, gw, dev)
    outs = U.Iterative_expand_matrix(scores, sx.to(dev), sy.to(dev), torch.tensor([0, gh, 0, gw], device=dev), rng, pos, height=gh, width=gw,
                                     iter_num=it, lower_bound=lb, return_nomatching=True)
    ref = oracle.iterative_expand_matrix(scores.cpu().numpy(), sx.numpy(), sy.numpy(), gh, gw, lower_bound=lb, iter_num=it)
    whole, core, avg, xs, ys, bound, nm = [o.cpu().numpy() for o in outs]
    assert np.array_equal(bound, ref[5]), f"{int((bound != ref[5]).any(-1).sum())} boxes differ"
    assert np.array_equal(nm, ref[6])
    assert (bound[..., 1] > bound[..., 0]).mean() > 0.2, "inputs must actually grow boxes"
    np.testing.assert_allclose(avg, ref[2], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(xs, ref[3], rtol=1e-6)
    np.testing.assert_allclose(ys, ref[4], rtol=1e-6)
    np.testing.assert_allclose(whole, ref[0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(core, ref[1], rtol=1e-4, atol=1e-6)


@pytest.fixture
def cpu_tie_break(monkeypatch):
    """The fixtures / the oracle hold the reference run on CPU tensors, where the unstable argsort of second_layer.py:169 / :230
    returns the FIRST of several equal minima; the library's default follows ATen's CUDA kernel (checked live below)."""
    from pats_b200 import layers as L

    monkeypatch.setattr(L, "MERGE_TIE_BREAK", "first")


def test_argsort_tie_rule_matches_the_live_cuda_op(dev):
    """kArgsortTie9 (csrc/regroup.cu) against torch.argsort on THIS GPU, on tie patterns like the merge produces; and the
    "first" mode against the first minimum."""
    import ctypes  # noqa: F401

    from pats_b200 import _lib

    g = torch.Generator().manual_seed(9)
    pool = torch.tensor([0.0, 0.0, 0.0, 1e-14, 8e-14, 100000.0, -9998.912109375, -9998.5, 0.25, 1.5], dtype=torch.float64)
    x = pool[torch.randint(0, len(pool), (60000, 9), generator=g)]
    x[:5000] = 0.0
    x[5000:10000] = torch.rand(5000, 9, generator=g, dtype=torch.float64)  # no ties
    xd = x.to(dev).contiguous()
    out = torch.empty(x.shape[0], dtype=torch.int32, device=dev)
    for tie_first in (0, 1):
        _lib.check(_lib.load().pats_argsort9_first_f64(xd.data_ptr(), x.shape[0], tie_first, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "argsort9")
        want = torch.argsort(xd.reshape(1, 200, 300, 9))[..., 0].reshape(-1) if tie_first == 0 else torch.from_numpy(np.argmin(x.numpy(), 1)).to(dev)
        assert torch.equal(out.long(), want.long()), f"tie_first={tie_first}: {int((out.long() != want.long()).sum())} of {x.shape[0]} rows differ"


@pytest.mark.parametrize("tag,merge_new", [("new", True), ("old", False)])
def test_merge_patches_golden(dev, tag, merge_new, cpu_tie_break):
    from pats_b200 import layers as L

    g = load_golden("merge")
    h, w = (int(v) for v in g["hw"])
    fn = L.merge_patches_new if merge_new else L.merge_patches_old
    for c in (0, 1):
        trust, nm2, sb = T(g[f"{tag}{c}_trust"], dev), T(g[f"{tag}{c}_nm2"], dev), T(g[f"{tag}{c}_sb_in"], dev)
        out, sb_out = fn(None, trust.shape[0], trust, [32 * h, 32 * w], T(g[f"{tag}{c}_nm1"], dev), nm2, sb)
        assert out.dtype == torch.bool
        assert np.array_equal(trust.cpu().numpy(), g[f"{tag}{c}_trust_after"]), "trust_score must be mutated in place like the reference"
        assert np.array_equal(nm2.cpu().numpy(), g[f"{tag}{c}_nm2_after"])
        assert np.array_equal(sb_out.cpu().numpy(), g[f"{tag}{c}_sb_out"])
        assert np.array_equal(out.cpu().numpy(), g[f"{tag}{c}_out"])


@pytest.mark.parametrize("merge_new", [True, False])
@pytest.mark.parametrize("h,w,frac", [(15, 20, 1.0), (15, 20, 0.6), (32, 32, 0.9), (3, 4, 1.0)])
def test_merge_patches_vs_oracle(dev, merge_new, h, w, frac, cpu_tie_break):
    from pats_b200 import layers as L

    g = torch.Generator().manual_seed(60 + h + int(merge_new))
    fn = L.merge_patches_new if merge_new else L.merge_patches_old
    sb = torch.zeros(1, h * w, 16, 9, dtype=torch.float64)
    sb_dev = sb.to(dev)
    for chunk in range(2):
        nm1 = ~(torch.rand(1, h * w, generator=g) < frac)
        P = int((~nm1).sum())
        trust = torch.rand(P, 144, generator=g) ** 3 * 1.2
        trust[torch.rand(P, 144, generator=g) < 0.05] = 1e-14
        nm2 = torch.rand(P, 144, generator=g) < 0.35
        ref_out, ref_sb, ref_trust, ref_nm2 = oracle.merge_patches(merge_new, trust.numpy(), [32 * h, 32 * w], nm1.numpy(), nm2.numpy(), sb.numpy())
        t_dev, f_dev = trust.to(dev), nm2.to(dev)
        out, sb_dev = fn(None, P, t_dev, [32 * h, 32 * w], nm1.to(dev), f_dev, sb_dev)
        assert np.array_equal(out.cpu().numpy(), ref_out), f"chunk {chunk}: {int((out.cpu().numpy() != ref_out).sum())} cells differ"
        assert np.array_equal(t_dev.cpu().numpy(), ref_trust) and np.array_equal(f_dev.cpu().numpy(), ref_nm2)
        assert np.array_equal(sb_dev.cpu().numpy(), ref_sb)
        sb = torch.from_numpy(ref_sb)


def test_get_result_golden_and_oracle(dev):
    from pats_b200 import utils as U

    g = load_golden("result")
    P = g["nm1"].shape[0]
    sc1 = np.repeat(g["sc1_first"][:, None, :], 2304, axis=1)
    choice = [torch.ones(1, dtype=torch.bool, device=dev), torch.ones(P, dtype=torch.bool, device=dev)]
    ml, mr = U.get_result(1, [T(g["nm0"], dev), T(g["nm1"], dev)], [T(g["pt0"], dev), T(g["pt1"].astype(np.float32), dev)],
                          [T(g["sc0"], dev), T(sc1, dev)], [[32, 15, 20], [2, 48, 48]], choice)
    assert ml.shape == g["matches_l"].shape
    assert np.array_equal(ml.cpu().numpy(), g["matches_l"]) and np.array_equal(mr.cpu().numpy(), g["matches_r"])
    # dense case: every fine cell of every window matched (K = 300 * 2304 = 691200 rows, the reference's worst case)
    gen = torch.Generator().manual_seed(70)
    nm0 = torch.zeros(1, 300, dtype=torch.bool)
    nm1 = torch.rand(300, 2304, generator=gen) < 0.02
    pt0 = torch.rand(1, 300, 2, generator=gen) * 15
    sc0 = torch.cat([torch.exp(torch.randn(1, 300, 1, generator=gen) * 0.4), torch.ones(1, 300, 1)], 2)
    pt1 = torch.rand(300, 2304, 2, generator=gen) * 48
    sc1 = sc0.reshape(300, 1, 2).repeat(1, 2304, 1)
    rl, rr = oracle.get_result([nm0.numpy(), nm1.numpy()], [pt0.numpy(), pt1.numpy()], [sc0.numpy(), sc1.numpy()], [[32, 15, 20], [2, 48, 48]])
    ml, mr = U.get_result(1, [nm0.to(dev), nm1.to(dev)], [pt0.to(dev), pt1.to(dev)], [sc0.to(dev), sc1.to(dev)], [[32, 15, 20], [2, 48, 48]], None)
    assert ml.shape == rl.shape and ml.shape[0] > 600000
    assert np.array_equal(ml.cpu().numpy(), rl) and np.array_equal(mr.cpu().numpy(), rr)
    # nothing matched at level 1
    ml, mr = U.get_result(1, [nm0.to(dev), torch.ones(300, 2304, dtype=torch.bool, device=dev)], [pt0.to(dev), pt1.to(dev)], [sc0.to(dev), sc1.to(dev)],
                          [[32, 15, 20], [2, 48, 48]], None)
    assert ml.shape == (0, 2) and mr.shape == (0, 2)


def test_third_compute_result_golden_and_oracle(dev):
    from pats_b200 import layers as L
    from pats_b200 import modules as M

    g = load_golden("third")
    sx = np.sqrt(g["scale"] + np.float32(1e-8))
    m0, m1, im = L.third_compute_result(T(np.exp(g["Z"]), dev), T(sx, dev), T(sx, dev), T(g["p_s"], dev), T(g["p_t"], dev))
    assert np.array_equal(im.cpu().numpy(), g["if_matching1"])
    assert np.array_equal(m0.cpu().numpy(), g["mkpts0_f"])
    np.testing.assert_allclose(m1.cpu().numpy(), g["mkpts1_f"], rtol=1e-5, atol=2e-5)
    r0, r1, _ = L.Compute_result(None, T(np.exp(g["Z"]), dev), 8, 5, T(sx, dev), T(sx, dev), T(g["p_s"], dev), T(g["p_t"], dev), dev)
    assert torch.equal(r0, m0) and torch.equal(r1, m1)
    # K = 4800 (one 640x480 pair's level-3 problems), against the oracle: same sequential f32 sums -> bit-exact
    gen = torch.Generator().manual_seed(80)
    K = 4800
    s = 3.0 * _planted(gen, 8, 8, 8, 1.0, 0.6).repeat(K // 8, 1, 1) + 0.3 * torch.randn(K, 64, 64, generator=gen)
    s = torch.cat([torch.cat([s, torch.full((K, 64, 1), -6.0)], 2), torch.full((K, 1, 65), -6.0)], 1)
    s[::7, :, -1] += 8.0
    scale = torch.exp((torch.rand(K, 1, 64, generator=gen) * 2 - 1) * math.log(16.0))
    scores = M.log_optimal_transport2(s.to(dev), 1.0, scale.to(dev), 100).exp()
    sxy = (scale + 1e-8).sqrt()
    p_s = torch.randint(0, 24, (K, 2), generator=gen) * 4
    p_t = torch.randint(0, 25, (K, 2), generator=gen) * 4
    m0, m1, im = L.third_compute_result(scores, sxy.to(dev), sxy.to(dev), p_s.to(dev), p_t.to(dev))
    o0, o1, oim = oracle.third_compute_result(scores.cpu().numpy(), sxy.numpy(), sxy.numpy(), p_s.numpy(), p_t.numpy())
    assert np.array_equal(im.cpu().numpy(), oim) and 0.05 < oim.mean() < 0.95
    assert np.array_equal(m0.cpu().numpy(), o0)
    assert np.array_equal(m1.cpu().numpy(), o1)


@pytest.mark.parametrize("tag,gh,gw,lb,it", [("L1", 15, 20, 1e-5, 15), ("L2", 12, 12, 1e-3, 8)])
def test_fused_est_position_equals_unfused(dev, tag, gh, gw, lb, it):
    """est_position from the log-domain plan (exp fused into the row staging) == masks + Iterative_expand_matrix on
    torch's scores.exp() -- and both equal the reference's golden outputs."""
    from pats_b200 import layers as L
    from pats_b200 import utils as U

    g = load_golden("expand")
    Z = T(g[tag + "_Z"], dev)
    sx, sy = T(g[tag + "_scalex"], dev), T(g[tag + "_scaley"], dev)
    trust, avg, xs, ys, nm1, nm2, core, bound = L.est_position(Z, sx, sy, gh, gw, it, lb, return_extra=True)
    rng, pos = _ranges_positions(gh, gw, dev)
    ref = U.Iterative_expand_matrix(Z.exp(), sx, sy, torch.tensor([0, gh, 0, gw], device=dev), rng, pos, height=gh, width=gw, iter_num=it,
                                    lower_bound=lb)
    assert torch.equal(bound, ref[5]) and np.array_equal(bound.cpu().numpy(), g[tag + "_bound"])
    assert torch.equal(trust, ref[0]) and torch.equal(core, ref[1]) and torch.equal(avg, ref[2])
    assert torch.equal(xs, ref[3]) and torch.equal(ys, ref[4])
    assert np.array_equal(nm1.cpu().numpy(), g[tag + "_nm1"]) and np.array_equal(nm2.cpu().numpy(), g[tag + "_nm2"])


def test_third_result_from_log_equals_exp_then_result(dev):
    from pats_b200 import layers as L

    g = load_golden("third")
    sx = T(np.sqrt(g["scale"] + np.float32(1e-8)), dev)
    Z = T(g["Z"], dev)
    a = L.third_result_from_log(Z, sx, sx, T(g["p_s"], dev), T(g["p_t"], dev))
    b = L.third_compute_result(Z.exp(), sx, sx, T(g["p_s"], dev), T(g["p_t"], dev))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert np.array_equal(a[2].cpu().numpy(), g["if_matching1"])


def test_grid_sample12_and_third_unfold_golden(dev):
    """a10 / a12: bit-exact against the reference statements (tests/golden/gathers.npz) and the oracle at real sizes."""
    from conftest import gathers_inputs
    from pats_b200 import layers as L

    g, maps, feat0, feat1 = gathers_inputs()
    out = L.grid_sample12([m.to(dev) for m in maps])
    assert np.array_equal(out.cpu().numpy(), g["gs_out"])
    kenc = T(g["un_kenc"], dev)
    o0 = L.third_unfold(feat0.to(dev), T(g["un_mk0"], dev), T(g["un_b"], dev), kenc, T(g["un_rubbish"], dev), T(g["un_mk0"], dev), False)
    o1 = L.third_unfold(feat1.to(dev), T(g["un_mk1"], dev), T(g["un_b"], dev), kenc, T(g["un_rubbish"], dev), T(g["un_mk0"], dev), True)
    assert np.array_equal(o0.cpu().numpy(), g["un_out0"]) and np.array_equal(o1.cpu().numpy(), g["un_out1"])
    with pytest.raises(IndexError):
        L.third_unfold(feat1.to(dev), torch.zeros(1, 2, device=dev), torch.zeros(1, device=dev), kenc, T(g["un_rubbish"], dev),
                       T(g["un_mk0"][:1], dev), True)
    # real sizes: 2P = 600 windows (a10), K = 4800 points over P = 300 windows (a12)
    gen = torch.Generator().manual_seed(90)
    big = [torch.randn(600, 64, 48, 48, generator=gen), torch.randn(600, 64, 24, 24, generator=gen), torch.randn(600, 128, 12, 12, generator=gen)]
    outb = L.grid_sample12([m.to(dev) for m in big]).cpu().numpy()
    sel = [0, 299, 599]
    ref = oracle.grid_sample12(big[0][sel].numpy(), big[1][sel].numpy(), big[2][sel].numpy())
    assert np.array_equal(outb[sel], ref)
    P, K = 300, 4800
    feat = torch.randn(P, 128, 52, 52, generator=gen)
    rub = torch.randn(P, 128, 144, generator=gen)
    kk = torch.randn(128, 64, generator=gen)
    cell = torch.randint(1, 11, (K, 2), generator=gen)
    mk0 = (cell * 4 + 2).float() * 2
    mk1 = (torch.rand(K, 2, generator=gen) * 80 + 8)
    b = torch.randint(0, P, (K,), generator=gen).float()
    o = L.third_unfold(feat.to(dev), mk1.to(dev), b.to(dev), kk.to(dev), rub.to(dev), mk0.to(dev), True).cpu().numpy()
    idx = torch.randperm(K, generator=gen)[:64]
    r = oracle.third_unfold(feat.numpy(), mk1[idx].numpy(), b[idx].numpy(), kk.numpy(), rub.numpy(), mk0[idx].numpy(), True)
    assert np.array_equal(o[idx.numpy()], r)


# ---- composite calls (second_layer.py:103-116, third_layer.py:158-167) with the plan hand-over ----------------------------
def _l2_inputs(b, n, seed, dev, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    s = (scale * torch.randn(b, n + 1, n + 1, generator=g)).to(dev)
    sx = torch.exp((torch.rand(b, n, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    sy = torch.exp((torch.rand(b, n, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    return s, sx, sy


@pytest.mark.parametrize("b,grid,outdoor,handover", [(7, 12, True, 1), (301, 12, True, 1), (301, 12, False, 0), (5, 4, True, 1), (3, 15, False, 1)])
def test_second_layer_match_equals_separate_calls(dev, b, grid, outdoor, handover):
    """One call == log_optimal_transport2, the in-place dustbin offsets of second_layer.py:108-112, est_position --
    bit for bit, with and without the early launch, on the 145 x 145 kernel and on shapes that take other kernels."""
    from pats_b200 import _lib, layers as Ly, modules as M

    n = grid * grid
    s, sx, sy = _l2_inputs(b, n, 100 + b + grid, dev)
    ns = (sx * sy).reshape(b, 1, n)
    one = torch.tensor(1.0, device=dev)
    Z = M.log_optimal_transport2(s, one, ns, 100)
    c = torch.log(one * (2 if outdoor else 3))
    Z[:, :, -1] += c
    Z[:, -1, :] += c
    ref = Ly.est_position(Z, sx, sy, grid, grid, 8, 1e-3, return_extra=True)
    lib = _lib.load()
    lib.pats_plan_handover(handover)
    try:
        for _ in range(3):  # repeated calls reuse the flag pool (epochs) and the self-cleaning scratch of est_position
            out = Ly.second_layer_match(s, one, ns, sx, sy, 100, outdoor, grid, return_extra=True)
            assert torch.equal(out[0], Z)
            for got, want in zip(out[1:], ref):
                assert torch.equal(got, want)
    finally:
        lib.pats_plan_handover(1)


def test_second_layer_match_vs_oracle(dev):
    from pats_b200 import layers as Ly

    b, n = 6, 144
    s, sx, sy = _l2_inputs(b, n, 77, dev)
    ns = (sx * sy).reshape(b, 1, n)
    Z, trust, avg, xs, ys, nm1, nm2, core, bound = Ly.second_layer_match(s, 1.0, ns, sx, sy, 100, True, 12, return_extra=True)
    Zo = oracle.log_optimal_transport2(s.cpu().numpy(), 1.0, ns.cpu().numpy(), 100)
    c = np.float32(np.log(np.float32(2.0)))
    Zo[:, :, -1] += c
    Zo[:, -1, :] += c
    np.testing.assert_allclose(Z.cpu().numpy(), Zo, atol=1e-4, rtol=0)
    # the integer results hang off the plan: compare them on the kernel's own plan (the oracle's plan differs by ~1e-6)
    w = oracle.iterative_expand_matrix(Z.exp().cpu().numpy(), sx.cpu().numpy(), sy.cpu().numpy(), 12, 12, 1e-3, 8)
    assert np.array_equal(bound.cpu().numpy(), w[5])
    o1, o2 = oracle.est_nomatching(Z.cpu().numpy(), n)
    assert np.array_equal(nm1.cpu().numpy(), o1) and np.array_equal(nm2.cpu().numpy(), o2)


@pytest.mark.parametrize("K,handover", [(5, 1), (4801, 1), (4801, 0)])
def test_third_layer_match_equals_separate_calls(dev, K, handover):
    from pats_b200 import _lib, layers as Ly, modules as M

    g = torch.Generator().manual_seed(K)
    s = (0.1 * torch.randn(K, 65, 65, generator=g)).to(dev)
    ns = torch.exp((torch.rand(K, 1, 64, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    sxy = (ns.reshape(K, 64) + 1e-8).sqrt()
    p_s = (torch.randint(0, 24, (K, 2), generator=g) * 4).to(dev)
    p_t = (torch.randint(0, 25, (K, 2), generator=g) * 4).to(dev)
    Z = M.log_optimal_transport2(s, 1.0, ns, 100)
    ref = Ly.third_result_from_log(Z, sxy, sxy, p_s, p_t)
    lib = _lib.load()
    lib.pats_plan_handover(handover)
    try:
        for _ in range(3):
            out = Ly.third_layer_match(s, 1.0, ns, sxy, sxy, p_s, p_t, 100)
            assert torch.equal(out[0], Z)
            for got, want in zip(out[1:], ref):
                assert torch.equal(got, want)
    finally:
        lib.pats_plan_handover(1)


def test_launch_chaining_on_and_off_give_identical_results(dev):
    """Programmatic-stream-serialization launches (griddepcontrol.wait at the top of every kernel) must not change anything:
    run a chain of dependent calls with chaining on and off and compare bit for bit."""
    from pats_b200 import _lib, layers as Ly, modules as M, utils as U

    g = torch.Generator().manual_seed(99)
    b, n = 40, 144
    s, sx, sy = _l2_inputs(b, n, 5, dev)
    ns = (sx * sy).reshape(b, 1, n)
    s1 = (0.1 * torch.randn(1, 300, 300, generator=g)).to(dev)
    ns1 = torch.exp((torch.rand(1, 1, 300, generator=g) * 2 - 1) * math.log(16.0)).to(dev)
    lib = _lib.load()

    def run():
        Z1 = M.log_optimal_transport(s1, 1.0, ns1, 100)
        e1 = Ly.est_position(Z1, ns1, ns1, 15, 20, 15, 1e-5, return_extra=True)
        out = Ly.second_layer_match(s, 1.0, ns, sx, sy, 100, True, 12, return_extra=True)
        nm_L1 = torch.zeros(1, 300, dtype=torch.bool, device=dev)
        nm_L1[0, b:] = True
        keep, sb = Ly.merge_patches_new(None, b, out[1].clone(), [480, 640], nm_L1, out[5].clone(), torch.zeros(1, 300, 16, 9, dtype=torch.float64, device=dev))
        torch.cuda.synchronize()
        return [Z1, *e1, *out, keep, sb]

    lib.pats_launch_chaining(0)
    try:
        ref = run()
    finally:
        lib.pats_launch_chaining(1)
    for _ in range(3):
        got = run()
        for a_, b_ in zip(got, ref):
            assert torch.equal(a_, b_)


def test_plan_handover_stress_sizes_streams_and_epochs(dev):
    """The per-stream flag pools (epochs, growth) and the self-cleaning est_position scratch under changing batch sizes on two
    streams: every composite result must equal the separately launched pair of calls, bit for bit."""
    from pats_b200 import layers as Ly, modules as M

    n = 144
    sizes = [300, 7, 301, 149, 2, 600, 296]
    data = {}
    for b in sizes:
        s, sx, sy = _l2_inputs(b, n, 900 + b, dev)
        ns = (sx * sy).reshape(b, 1, n)
        Z = M.log_optimal_transport2(s, 1.0, ns, 30)
        c = torch.log(torch.tensor(2.0, device=dev))
        Z[:, :, -1] += c
        Z[:, -1, :] += c
        data[b] = (s, sx, sy, ns, Z, Ly.est_position(Z, sx, sy, 12, 12, 8, 1e-3, return_extra=True))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    results = []
    for rep in range(4):
        for i, b in enumerate(sizes):
            s, sx, sy, ns, Z, ref = data[b]
            with torch.cuda.stream(streams[(i + rep) % 2]):
                results.append((b, Ly.second_layer_match(s, 1.0, ns, sx, sy, 30, True, 12, return_extra=True)))
    torch.cuda.synchronize()
    for b, out in results:
        _, _, _, _, Z, ref = data[b]
        assert torch.equal(out[0], Z), b
        for got, want in zip(out[1:], ref):
            assert torch.equal(got, want), b
