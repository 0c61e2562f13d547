"""GPU parity tests of the Sinkhorn / OT kernels, called through the reference-named wrappers (C ABI).

Bar (BASELINE.json north_star): OT scores within 1e-4 (f32, log domain) of the reference, row/column
argmax (match indices) identical.  Checked against (1) the committed outputs of the unmodified reference
(tests/golden/ot.npz), (2) the CPU oracle on seeded inputs, (3) size-independent properties at full size.
"""
import math

import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def M():
    from pats_b200 import modules

    return modules


@pytest.fixture(scope="module")
def lib():
    from pats_b200 import _lib

    return _lib.load()


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def areas(g, b, n, span):
    return torch.exp((torch.rand(b, 1, n, generator=g) * 2 - 1) * math.log(span))


def assert_plan_equal(out, ref, tol=TOL):
    assert out.shape == ref.shape
    err = np.abs(out - ref).max()
    assert err <= tol, f"max|err| = {err}"
    assert (out.argmax(2) == ref.argmax(2)).all(), "row argmax (match index) differs"
    assert (out.argmax(1) == ref.argmax(1)).all(), "column argmax (match index) differs"


# ---- (1) golden vectors of the unmodified reference ------------------------------------------------------
def test_golden_log_sinkhorn_iterations(M, dev):
    g = load_golden("ot")
    for tag, it in (("a1_out_it100", 100), ("a1_out_it3", 3)):
        out = M.log_sinkhorn_iterations(T(g["a1_Z"], dev), T(g["a1_log_mu"], dev), T(g["a1_log_nu"], dev), it).cpu().numpy()
        np.testing.assert_allclose(out, g[tag], atol=TOL, rtol=0)


@pytest.mark.parametrize("tag", ["a2_small", "a2_rect", "a2_L1", "a2_L1_peaked"])
def test_golden_log_optimal_transport(M, dev, tag):
    g = load_golden("ot")
    out = M.log_optimal_transport(T(g[tag + "_scores"], dev), torch.tensor(float(g[tag + "_alpha"]), device=dev), T(g[tag + "_ns"], dev), 100)
    assert out.is_contiguous() and out.dtype == torch.float32
    assert_plan_equal(out.cpu().numpy(), g[tag + "_out"])


def test_golden_single_iteration(M, dev):
    g = load_golden("ot")
    out = M.log_optimal_transport(T(g["a2_small_scores"], dev), float(g["a2_small_alpha"]), T(g["a2_small_ns"], dev), 1)
    np.testing.assert_allclose(out.cpu().numpy(), g["a2_small_out_it1"], atol=TOL, rtol=0)


@pytest.mark.parametrize("tag", ["a3_L2", "a3_L3", "a3_L3_wide", "a3_rect"])
def test_golden_log_optimal_transport2(M, dev, tag):
    g = load_golden("ot")
    s = T(g[tag + "_scores"], dev)
    s0 = s.clone()
    out = M.log_optimal_transport2(s, torch.tensor(1.0, device=dev), T(g[tag + "_ns"], dev), 100)
    assert torch.equal(s, s0), "input must not be modified"
    assert out.data_ptr() != s.data_ptr()
    assert_plan_equal(out.cpu().numpy(), g[tag + "_out"])


# ---- (2) CPU oracle on seeded inputs -------------------------------------------------------------------
@pytest.mark.parametrize("b,m,n,scale,span", [
    (64, 65, 65, 0.1, 16.0),     # level 3 shape, warp kernel
    (37, 65, 65, 1.5, 16.0),     # ragged batch (not a multiple of the 4 problems per CTA), wider scores
    (12, 145, 145, 0.1, 256.0),  # level 2 shape, CTA kernel
    (5, 145, 145, 1.0, 256.0),
    (9, 33, 60, 0.5, 4.0),       # non-square, padded tile
    (7, 72, 20, 0.5, 4.0),       # capacity edge of the warp kernel (72 rows)
    (3, 2, 2, 1.0, 2.0),         # smallest legal transport2 problem
    (5, 17, 31, 0.3, 8.0),       # tiny kernel
    (4, 100, 160, 0.4, 8.0),     # CTA kernel, capacity edge (160 columns)
    (3, 160, 73, 0.4, 8.0),      # CTA kernel, capacity edge (160 rows)
    (2, 161, 90, 0.4, 8.0),      # first shape that needs the cluster kernel
    (3, 301, 301, 0.1, 16.0),    # level-1 size with the dustbin in place, 8-CTA cluster (320 x 320 capacity)
    (2, 320, 320, 0.3, 16.0),    # capacity edge of the 320 cluster
    (2, 321, 200, 0.3, 16.0),    # needs the 512 cluster
    (1, 512, 512, 0.2, 16.0),    # capacity edge of the 512 cluster
    (5, 200, 500, 0.2, 16.0),    # ragged: most CTAs of the cluster own no valid rows
    (1, 513, 40, 0.3, 4.0),      # first shape that needs the generic log-domain kernel
])
def test_transport2_vs_oracle(M, lib, dev, b, m, n, scale, span):
    g = torch.Generator().manual_seed(1000 + b * 7 + m)
    s = scale * torch.randn(b, m, n, generator=g)
    ns = areas(g, b, n - 1, span)
    out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), 100).cpu().numpy()
    ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), 100)
    assert_plan_equal(out, ref)


@pytest.mark.parametrize("iters", [0, 1, 2, 100])
def test_level3_direct_start_and_log_start_agree_with_the_oracle(M, dev, iters):
    """The 65 x 65 kernel starts directly on exp(Z) when every |z| <= 12 and takes the log-domain first iteration otherwise
    (decided per problem).  Mix both kinds in one batch, at the threshold, for the iteration counts that take different
    code paths (0, 1: no / one scaling pass; 2; 100)."""
    g = torch.Generator().manual_seed(4242 + iters)
    b = 96
    s = 0.5 * torch.randn(b, 65, 65, generator=g)
    s[0::4, 3, 7] = 12.0          # exactly at the bound: direct
    s[1::4, 11, 60] = 12.5        # just beyond: log-domain start
    s[2::4, 64, 5] = -13.0        # dustbin row entry beyond the bound
    s[3::4] = s[3::4].clamp(-11.9, 11.9)
    ns = areas(g, b, 64, 16.0)
    out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
    ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
    assert_plan_equal(out, ref)


@pytest.mark.parametrize("b,m,n,alpha", [(1, 300, 300, 1.0), (3, 64, 64, 0.7), (2, 20, 33, 0.0), (2, 144, 144, 2.0), (1, 160, 100, 1.0)])
def test_transport_vs_oracle(M, dev, b, m, n, alpha):
    g = torch.Generator().manual_seed(2000 + m)
    s = 0.1 * torch.randn(b, m, n, generator=g)
    ns = areas(g, b, n, 16.0)
    out = M.log_optimal_transport(s.to(dev), torch.tensor(alpha, device=dev), ns.to(dev), 100)
    assert out.shape == (b, m + 1, n + 1)
    ref = oracle.log_optimal_transport(s.numpy(), alpha, ns.numpy(), 100)
    assert_plan_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("iters", [0, 1, 2, 3, 9, 17])
def test_iteration_counts(M, dev, iters):
    g = torch.Generator().manual_seed(3000 + iters)
    for (b, m, n) in ((6, 65, 65), (3, 145, 145), (2, 301, 301)):
        s = 0.3 * torch.randn(b, m, n, generator=g)
        ns = areas(g, b, n - 1, 16.0)
        out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
        ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
        np.testing.assert_allclose(out, ref, atol=TOL, rtol=0)


def test_raw_sinkhorn_vs_oracle(M, dev):
    g = torch.Generator().manual_seed(4000)
    for (b, m, n) in ((5, 65, 65), (3, 40, 150), (2, 200, 180)):
        Z = 0.5 * torch.randn(b, m, n, generator=g)
        lmu = torch.log_softmax(torch.randn(b, m, generator=g), 1)
        lnu = torch.log_softmax(torch.randn(b, n, generator=g), 1)
        out = M.log_sinkhorn_iterations(Z.to(dev), lmu.to(dev), lnu.to(dev), 50).cpu().numpy()
        ref = oracle.log_sinkhorn_iterations(Z.numpy(), lmu.numpy(), lnu.numpy(), 50)
        np.testing.assert_allclose(out, ref, atol=TOL, rtol=0)


def test_generic_kernel_matches_register_kernels(M, lib, dev):
    """The log-domain kernel (fallback / large shapes) and the register kernels agree."""
    g = torch.Generator().manual_seed(5000)
    for (b, m, n) in ((8, 65, 65), (3, 145, 145), (2, 301, 301)):
        s = (0.2 * torch.randn(b, m, n, generator=g)).to(dev)
        ns = areas(g, b, n - 1, 16.0).to(dev)
        fast = M.log_optimal_transport2(s, 1.0, ns, 100)
        lib.pats_sinkhorn_force_generic(1)
        try:
            assert lib.pats_sinkhorn_kernel_kind(m, n) == 2
            slow = M.log_optimal_transport2(s, 1.0, ns, 100)
        finally:
            lib.pats_sinkhorn_force_generic(0)
        assert (fast - slow).abs().max().item() <= TOL


def test_extreme_dynamic_range_takes_the_log_domain_fallback(M, lib, dev):
    """Scores with a huge spread drive the scaling form out of its safe range; those problems must be
    re-solved in the log domain and still match the oracle."""
    g = torch.Generator().manual_seed(6000)
    b, m, n = 6, 65, 65
    s = 0.1 * torch.randn(b, m, n, generator=g)
    s[0] *= 400.0       # |Z| up to ~150: far outside anything the exp-domain form can hold
    s[3, :, :5] -= 90.0
    ns = areas(g, b, n - 1, 16.0)
    lib.pats_sinkhorn_fallback_count(1)
    out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), 100).cpu().numpy()
    n_fb = lib.pats_sinkhorn_fallback_count(1)
    ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), 100)
    assert np.isfinite(out).all()
    # large-magnitude entries: relative tolerance on top of the absolute one
    np.testing.assert_allclose(out, ref, atol=TOL, rtol=2e-6)
    assert n_fb >= 1, "expected at least one problem to use the fallback"
    assert n_fb < b, "well-conditioned problems must stay on the fast path"


def test_all_three_65x65_kernels_agree(M, lib, dev):
    """65 x 65 normally runs on the two-warps-per-problem kernel; the one-warp 65 x 65 kernel (mode 2) and the padded
    72 x 68 warp kernel (mode 1) must agree with it and with the oracle on the same problems, including an odd batch,
    a wide score range that needs the fallback, and every iteration-count edge."""
    g = torch.Generator().manual_seed(6200)
    b = 21
    s = (0.3 * torch.randn(b, 65, 65, generator=g))
    s[4] *= 300.0
    ns = areas(g, b, 64, 16.0)
    for iters in (100, 0, 1, 2, 9):
        ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
        for mode in (0, 1, 2, 3):
            lib.pats_sinkhorn_disable_w65(mode)
            try:
                out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
            finally:
                lib.pats_sinkhorn_disable_w65(0)
            keep65 = [i for i in range(b) if i != 4]
            np.testing.assert_allclose(out[keep65], ref[keep65], atol=TOL, rtol=0, err_msg=f"mode {mode} iters {iters}")
            np.testing.assert_allclose(out[4], ref[4], atol=TOL, rtol=2e-5, err_msg=f"mode {mode} iters {iters} (fallback problem)")
            if iters == 100:
                keep = [i for i in range(b) if i != 4]
                assert (out[keep].argmax(2) == ref[keep].argmax(2)).all() and (out[keep].argmax(1) == ref[keep].argmax(1)).all()
    # augmenting transport with m = n = 64 and the raw iteration also land on the 65 x 65 kernel
    s64 = 0.2 * torch.randn(5, 64, 64, generator=g)
    ns64 = areas(g, 5, 64, 16.0)
    out = M.log_optimal_transport(s64.to(dev), 0.8, ns64.to(dev), 100).cpu().numpy()
    assert_plan_equal(out, oracle.log_optimal_transport(s64.numpy(), 0.8, ns64.numpy(), 100))
    Z = 0.5 * torch.randn(4, 65, 65, generator=g)
    lmu = torch.log_softmax(torch.randn(4, 65, generator=g), 1)
    lnu = torch.log_softmax(torch.randn(4, 65, generator=g), 1)
    out = M.log_sinkhorn_iterations(Z.to(dev), lmu.to(dev), lnu.to(dev), 37).cpu().numpy()
    np.testing.assert_allclose(out, oracle.log_sinkhorn_iterations(Z.numpy(), lmu.numpy(), lnu.numpy(), 37), atol=TOL, rtol=0)


def test_kernel_variants_145_and_cluster(M, lib, dev):
    """145 x 145: dedicated 9-warp kernel vs the padded 160 x 160 CTA kernel; 301 x 301: every cluster shape (auto = 10 CTAs x 256
    for small batches, 4 x 512, 8 x 256 one-hop, 8 x 256, 10 x 256) and the auto rule's large-batch side."""
    g = torch.Generator().manual_seed(6300)
    s = 0.3 * torch.randn(7, 145, 145, generator=g)
    s[2] *= 250.0  # one problem that needs the log-domain fallback
    ns = areas(g, 7, 144, 256.0)
    for iters in (100, 0, 1, 2, 10):
        ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
        for off in (0, 1, 2):
            lib.pats_sinkhorn_disable_c145(off)
            try:
                out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
            finally:
                lib.pats_sinkhorn_disable_c145(0)
            keep = [i for i in range(7) if i != 2]
            np.testing.assert_allclose(out[keep], ref[keep], atol=TOL, rtol=0, err_msg=f"c145 off={off} iters={iters}")
            # the extreme problem (|Z| in the hundreds) runs the f32 log-domain fallback: f32 rounding at that magnitude
            np.testing.assert_allclose(out[2], ref[2], atol=TOL, rtol=2e-5, err_msg=f"c145 off={off} iters={iters} (fallback problem)")
    s = 0.1 * torch.randn(2, 300, 300, generator=g)
    ns = areas(g, 2, 300, 16.0)
    ref = oracle.log_optimal_transport(s.numpy(), 1.0, ns.numpy(), 100)
    for v in (0, 1, 2, 3, 4):
        lib.pats_sinkhorn_cluster_variant(v)
        try:
            out = M.log_optimal_transport(s.to(dev), 1.0, ns.to(dev), 100).cpu().numpy()
        finally:
            lib.pats_sinkhorn_cluster_variant(0)
        assert_plan_equal(out, ref)
    # b > 8 takes the portable 8-CTA clusters under the auto rule; all problems must still be solved (repeat the two plans)
    s9, ns9 = s.repeat(5, 1, 1)[:9].contiguous(), ns.repeat(5, 1, 1)[:9].contiguous()
    out9 = M.log_optimal_transport(s9.to(dev), 1.0, ns9.to(dev), 100).cpu().numpy()
    for i in range(9):
        assert_plan_equal(out9[i:i + 1], ref[i % 2:i % 2 + 1])


def test_cluster_kernel_fallback(M, lib, dev):
    """Same as above for the 8-CTA cluster kernel: the cluster must agree on the verdict and rank 0 re-solves."""
    g = torch.Generator().manual_seed(6100)
    b, m, n = 3, 301, 301
    s = 0.1 * torch.randn(b, m, n, generator=g)
    s[1] *= 300.0
    ns = areas(g, b, n - 1, 16.0)
    lib.pats_sinkhorn_fallback_count(1)
    out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), 30).cpu().numpy()
    n_fb = lib.pats_sinkhorn_fallback_count(1)
    ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), 30)
    np.testing.assert_allclose(out, ref, atol=TOL, rtol=2e-6)
    assert n_fb == 1


def _peaked(g, b, gw, peak, floor):
    """[b, gw*gw+1, gw*gw+1]: affinity of a random affine warp of a gw x gw grid + dustbin row / column (plans a trained matcher produces)."""
    n = gw * gw
    ys, xs = torch.meshgrid(torch.arange(gw).float(), torch.arange(gw).float(), indexing="ij")
    src = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
    A = torch.eye(2)[None] * (0.6 + 0.8 * torch.rand(b, 1, 1, generator=g)) + 0.1 * torch.randn(b, 2, 2, generator=g)
    w = (src - gw / 2.0) @ A.transpose(1, 2) + gw / 2.0 + torch.randn(b, 1, 2, generator=g) * (0.12 * gw)
    d2 = ((w[:, :, None, :] - src[None, None, :, :]) ** 2).sum(-1)
    out = torch.empty(b, n + 1, n + 1)
    out[:, :n, :n] = (peak - d2 / 1.5 + 0.3 * torch.randn(b, n, n, generator=g)).clamp_min(floor)
    out[:, n, :] = -1.0 + 0.3 * torch.randn(b, n + 1, generator=g)
    out[:, :, n] = -1.0 + 0.3 * torch.randn(b, n + 1, generator=g)
    return out


@pytest.mark.parametrize("gw", [8, 12])
@pytest.mark.parametrize("kind", ["peaked_direct", "peaked_log_start", "diffuse", "ill_conditioned", "few_iterations"])
def test_fixed_point_exit_is_bit_identical(M, lib, dev, kind, gw):
    """The 65 x 65 kernel leaves its loop once a whole iteration left every beta bit-identical (pats_sinkhorn_fixed_point_exit).
    gw = 8: the 65 x 65 level-3 kernel, which has the exit; gw = 12: the 145 x 145 level-2 kernel, which does not (the switch must be a no-op there).
    The reference runs a fixed 100 iterations (models/modules.py:139-142), so the exit is only legitimate if the result is
    IDENTICAL, bit for bit, to running all of them -- on every kind of input, including problems that never converge, problems
    that start in the log domain and problems that end in the log-domain fallback."""
    g = torch.Generator().manual_seed(77 + gw)
    n = gw * gw
    b, iters = (1500 if gw == 8 else 320), 100
    if kind == "peaked_direct":
        s = _peaked(g, b, gw, 7.0, -10.0)         # |z| <= 12: direct start (65 x 65 kernel)
    elif kind == "peaked_log_start":
        s = _peaked(g, b, gw, 14.0, -25.0)        # log-domain first iteration
    elif kind == "diffuse":
        s = 0.1 * torch.randn(b, n + 1, n + 1, generator=g)
    elif kind == "ill_conditioned":
        s = 40.0 * torch.randn(b, n + 1, n + 1, generator=g)   # most problems end in the fallback
    else:
        s, iters = _peaked(g, b, gw, 7.0, -10.0), 7
    ns = areas(g, b, n, 4.0).to(dev)
    s = s.to(dev)
    one = torch.tensor(1.0, device=dev)
    try:
        lib.pats_sinkhorn_fixed_point_exit(0)
        lib.pats_sinkhorn_iterations_skipped(1)
        full = M.log_optimal_transport2(s, one, ns, iters)
        torch.cuda.synchronize()
        assert lib.pats_sinkhorn_iterations_skipped(1) == 0, "exit disabled, yet iterations were skipped"
        lib.pats_sinkhorn_fixed_point_exit(1)
        fast = M.log_optimal_transport2(s, one, ns, iters)
        torch.cuda.synchronize()
        skipped = lib.pats_sinkhorn_iterations_skipped(1)
    finally:
        lib.pats_sinkhorn_fixed_point_exit(1)
    same = torch.equal(full, fast) or bool(((full == fast) | (full.isnan() & fast.isnan())).all())
    assert same, f"{kind}: {int((full != fast).sum())} entries differ between the early exit and the full {iters} iterations"
    if kind in ("peaked_direct", "diffuse") and gw == 8:  # the 145 x 145 kernel has no exit (measured: DESIGN.md section 8); its cases pin that
        assert skipped > 0, f"{kind}: the exit never fired (it is not exercised by this test)"
    print(f"{kind}: {skipped} of {b * iters} problem-iterations skipped")


@pytest.mark.parametrize("b", [1, 4, 5, 63, 1187, 4800])
def test_bulk_staging_is_bit_identical(M, lib, dev, b):
    """pats_sinkhorn_bulk_staging(1): persistent CTAs, each 65 x 65 problem staged by one cp.async.bulk of its 16-byte aligned
    superset and written back by one bulk store.  Same arithmetic, so the plans must be bit-identical to the direct kernel --
    for batch sizes that end inside / outside a 16-byte boundary (b % 4), more problems than CTAs, ill-conditioned problems
    (in-kernel fallback), raw marginals, and through the composite call (plan hand-over flags raised after the bulk store)."""
    from pats_b200 import layers as Ly

    g = torch.Generator().manual_seed(500 + b)
    s = _peaked(g, b, 8, 7.0, -10.0)
    if b >= 63:
        s[7] *= 40.0  # one problem that ends in the log-domain fallback
    s = s.to(dev)
    ns = areas(g, b, 64, 4.0).to(dev)
    one = torch.tensor(1.0, device=dev)
    lmu = torch.log_softmax(torch.randn(b, 65, generator=g), 1).to(dev)
    lnu = torch.log_softmax(torch.randn(b, 65, generator=g), 1).to(dev)
    sxy = (ns.reshape(b, 64) + 1e-8).sqrt()
    ps = (torch.randint(0, 24, (b, 2), generator=g) * 4).to(dev)
    outs = {}
    try:
        for mode in (0, 1):
            lib.pats_sinkhorn_bulk_staging(mode)
            outs[mode] = (M.log_optimal_transport2(s, one, ns, 100), M.log_sinkhorn_iterations(s, lmu, lnu, 20),
                          Ly.third_layer_match(s, 1.0, ns, sxy, sxy, ps, ps, 100))
            torch.cuda.synchronize()
    finally:
        lib.pats_sinkhorn_bulk_staging(1)
    eq = lambda x, y: bool(((x == y) | (x.isnan() & y.isnan())).all()) if x.is_floating_point() else torch.equal(x, y)  # noqa: E731
    assert eq(outs[0][0], outs[1][0]), "log_optimal_transport2 differs with bulk staging"
    assert eq(outs[0][1], outs[1][1]), "log_sinkhorn_iterations differs with bulk staging"
    for x, y in zip(outs[0][2], outs[1][2]):
        assert eq(x, y), "third_layer_match differs with bulk staging"


def test_unaligned_views_take_the_direct_kernel_and_agree(M, dev):
    """A [1:] view of a [b,65,65] tensor starts 16 900 B into the allocation (4-byte phase): the bulk-copy staging needs 16-byte
    aligned tensors, so the dispatch falls back to the direct loads.  Same bits either way."""
    g = torch.Generator().manual_seed(91)
    s = _peaked(g, 41, 8, 7.0, -10.0).to(dev)
    ns = areas(g, 41, 64, 4.0).to(dev)
    one = torch.tensor(1.0, device=dev)
    view = s[1:]
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    a = M.log_optimal_transport2(view, one, ns[1:], 100)
    b = M.log_optimal_transport2(view.clone(), one, ns[1:].clone(), 100)
    assert torch.equal(a, b)


def test_empty_batch(M, dev):
    out = M.log_optimal_transport2(torch.zeros(0, 65, 65, device=dev), 1.0, torch.zeros(0, 1, 64, device=dev), 100)
    assert out.shape == (0, 65, 65)
    out = M.log_optimal_transport(torch.zeros(0, 30, 30, device=dev), 1.0, torch.zeros(0, 1, 30, device=dev), 100)
    assert out.shape == (0, 31, 31)


def test_rejects_cpu_tensors(M):
    with pytest.raises(RuntimeError, match="CUDA-only"):
        M.log_optimal_transport2(torch.zeros(1, 5, 5), 1.0, torch.ones(1, 1, 4), 10)


# ---- (3) full-size, size-independent properties --------------------------------------------------------------
def _torch_lse_sinkhorn(Z, lmu, lnu, iters):
    u, v = torch.zeros_like(lmu), torch.zeros_like(lnu)
    for _ in range(iters):
        u = lmu - torch.logsumexp(Z + v[:, None, :], 2)
        v = lnu - torch.logsumexp(Z + u[:, :, None], 1)
    return Z + u[:, :, None] + v[:, None, :]


@pytest.mark.parametrize("b,m,n,span", [(4800, 65, 65, 16.0), (300, 145, 145, 256.0)])
def test_full_size_marginals_and_subset_parity(M, dev, b, m, n, span):
    """BASELINE sizes (K=4800 level-3 problems / P=300 level-2 problems of one 640x480 pair):
    the column marginals are met exactly after the last v-update, and a random subset equals the oracle."""
    g = torch.Generator().manual_seed(7000 + m)
    s = (0.1 * torch.randn(b, m, n, generator=g)).to(dev)
    ns = areas(g, b, n - 1, span).to(dev)
    out = M.log_optimal_transport2(s, 1.0, ns, 100)
    ms = float(m - 1)
    tot = ms + ns.sum(2)                                   # [b,1]
    nu = torch.cat([ns.reshape(b, -1), torch.full((b, 1), ms, device=dev)], 1)   # column masses (x (m+sum ns) scaling)
    col = torch.logsumexp(out, 1)                          # log of column sums of the scaled plan
    assert (col - nu.log()).abs().max().item() < 2e-4
    mu = torch.cat([torch.ones(b, m - 1, device=dev), ns.sum(2)], 1)
    row = torch.logsumexp(out, 2)
    assert (row - mu.log()).abs().max().item() < 0.2       # rows only approximately (not converged exactly)
    assert abs(float(torch.exp(out).sum((1, 2)).mean() / tot.mean()) - 1.0) < 1e-3
    idx = torch.randperm(b, generator=g)[:24]
    ref = oracle.log_optimal_transport2(s[idx].cpu().numpy(), 1.0, ns[idx].cpu().numpy(), 100)
    assert_plan_equal(out[idx].cpu().numpy(), ref)


@pytest.mark.parametrize("N,iters", [(1536, 100), (1024, 100), (4096, 40)])  # 4096: BASELINE.json's stress size (four warps per row)
def test_large_plan_grid_kernel_vs_torch(M, lib, dev, N, iters):
    """BASELINE.json's synthetic kernel sizes (N=1536) and the 1024x1024-pair coarse plan (1024 -> 1025):
    grid-cooperative streaming kernel against a torch fp32 logsumexp restatement on the same device."""
    g = torch.Generator().manual_seed(8000 + N)
    b = 2
    s = (0.1 * torch.randn(b, N, N, generator=g)).to(dev)
    ns = areas(g, b, N, 16.0).to(dev)
    alpha = torch.tensor(1.0, device=dev)
    assert lib.pats_sinkhorn_kernel_kind(N + 1, N + 1) == 4
    out = M.log_optimal_transport(s, alpha, ns, iters)
    Z = torch.cat([torch.cat([s, alpha.expand(b, N, 1)], 2), alpha.expand(b, 1, N + 1)], 1)
    nsum = ns.sum(2).reshape(b)
    norm = -(N + nsum).log()
    lnu = torch.cat([ns.reshape(b, N).log() + norm[:, None], (math.log(N) + norm)[:, None]], 1)
    lmu = torch.cat([norm[:, None].expand(b, N), (nsum.log() + norm)[:, None]], 1)
    ref = _torch_lse_sinkhorn(Z, lmu, lnu, iters) - norm[:, None, None]
    assert (out - ref).abs().max().item() <= TOL
    assert torch.equal(out.argmax(2), ref.argmax(2))


def test_4096_column_kernels_vs_torch_in_every_mode(M, lib, dev):
    """4096 core columns (BASELINE.json's stress size) has its own kernel -- a row split over four warps, sinkhorn_gridq_kernel -- and
    the one-warp-per-row kernel behind pats_sinkhorn_grid_variant(2): both against the torch float32 logsumexp restatement in all
    three modes (dustbin synthesised, b = 2, 40 iterations; dustbin in memory with 37 and 601 rows at row stride 4097; raw marginals
    at 0 / 1 / 2 / 7 iterations).  The checks are tools/ab_grid4096.py's (profiles/r02_ab_grid4096.json is its record)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ab_grid4096 as A

    out = {"parity": [], "timing": []}
    try:
        A.parity(lib, M, dev, 4096, out)
    finally:
        lib.pats_sinkhorn_grid_variant(0)
    assert len(out["parity"]) == 14
    for r in out["parity"]:
        assert r["finite"] and r["argmax_equal"] and r["max_abs_diff_vs_torch_lse"] <= TOL, r


# ---- grid-cooperative streaming kernel (plans beyond 512 x 512) ---------------------------------------------
@pytest.mark.parametrize("mode,b,m,n,iters", [
    ("ot", 1, 1024, 1024, 100),    # level-1 plan of a 1024 x 1024 pair (1025 x 1025), all SMs on one problem
    ("ot", 3, 600, 513, 30),       # core width 513 -> 32 columns per lane, ragged
    ("ot2", 2, 1537, 1537, 30),    # BASELINE.json's N = 1536 (48 columns per lane)
    ("ot2", 2, 530, 513, 40),      # core width 512 -> 16 columns per lane
    ("ot2", 1, 700, 2049, 12),     # core width 2048 -> 64 columns per lane, 8 warps
    ("ot2", 1, 520, 2300, 8),      # > 2048 columns: exponentials recomputed instead of kept
    ("raw", 2, 513, 40, 25),       # tall and narrow: M alone exceeds the cluster kernel
    ("raw", 150, 520, 30, 10),     # more problems than SMs: one CTA per problem, problems looped
])
def test_grid_kernel_vs_oracle(M, lib, dev, mode, b, m, n, iters):
    g = torch.Generator().manual_seed(9000 + m + n)
    if mode == "ot":
        s = 0.1 * torch.randn(b, m, n, generator=g)
        ns = areas(g, b, n, 16.0)
        assert lib.pats_sinkhorn_kernel_kind(m + 1, n + 1) == 4
        out = M.log_optimal_transport(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
        ref = oracle.log_optimal_transport(s.numpy(), 1.0, ns.numpy(), iters)
    elif mode == "ot2":
        s = 0.1 * torch.randn(b, m, n, generator=g)
        ns = areas(g, b, n - 1, 16.0)
        assert lib.pats_sinkhorn_kernel_kind(m, n) == 4
        out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), iters).cpu().numpy()
        ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
    else:
        s = 0.5 * torch.randn(b, m, n, generator=g)
        lmu = torch.log_softmax(torch.randn(b, m, generator=g), 1)
        lnu = torch.log_softmax(torch.randn(b, n, generator=g), 1)
        assert lib.pats_sinkhorn_kernel_kind(m, n) == 4
        out = M.log_sinkhorn_iterations(s.to(dev), lmu.to(dev), lnu.to(dev), iters).cpu().numpy()
        ref = oracle.log_sinkhorn_iterations(s.numpy(), lmu.numpy(), lnu.numpy(), iters)
    assert_plan_equal(out, ref)


def test_grid_kernel_cta_split_and_iteration_counts(M, lib, dev):
    """Any split of the rows over CTAs gives the same plan (to rounding), for every iteration-count branch."""
    g = torch.Generator().manual_seed(9500)
    b, m, n = 2, 640, 600
    s = 0.1 * torch.randn(b, m, n, generator=g)
    ns = areas(g, b, n - 1, 16.0)
    sd, nsd = s.to(dev), ns.to(dev)
    for iters in (0, 1, 2, 9):
        ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), iters)
        for G in (0, 1, 3, 40, 74):
            lib.pats_sinkhorn_grid_ctas_per_problem(G)
            try:
                out = M.log_optimal_transport2(sd, 1.0, nsd, iters).cpu().numpy()
            finally:
                lib.pats_sinkhorn_grid_ctas_per_problem(0)
            np.testing.assert_allclose(out, ref, atol=TOL, rtol=0, err_msg=f"iters={iters} G={G}")


def test_grid_kernel_flags_extreme_problems_for_the_log_domain_kernel(M, lib, dev):
    g = torch.Generator().manual_seed(9600)
    b, m, n = 3, 560, 530
    s = 0.1 * torch.randn(b, m, n, generator=g)
    s[1] *= 300.0
    ns = areas(g, b, n - 1, 16.0)
    lib.pats_sinkhorn_fallback_count(1)
    out = M.log_optimal_transport2(s.to(dev), 1.0, ns.to(dev), 20).cpu().numpy()
    n_fb = lib.pats_sinkhorn_fallback_count(1)
    ref = oracle.log_optimal_transport2(s.numpy(), 1.0, ns.numpy(), 20)
    assert np.isfinite(out).all()
    np.testing.assert_allclose(out[[0, 2]], ref[[0, 2]], atol=TOL, rtol=0)
    np.testing.assert_allclose(out[1], ref[1], atol=TOL, rtol=2e-5)
    assert n_fb == 1
