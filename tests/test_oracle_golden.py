"""Pins the CPU oracle (oracle/pats_oracle.c) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only.

Tolerances: integer / byte / index results bit-exact; f32 OT outputs within 1e-4 (north_star);
other f32 results within a few ulp-scale relative error (sum order differs from ATen's).
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden

OT_TOL = 1e-4


def test_a1_log_sinkhorn_iterations():
    g = load_golden("ot")
    for tag, it in (("a1_out_it100", 100), ("a1_out_it3", 3)):
        out = oracle.log_sinkhorn_iterations(g["a1_Z"], g["a1_log_mu"], g["a1_log_nu"], it)
        np.testing.assert_allclose(out, g[tag], atol=OT_TOL, rtol=0)


@pytest.mark.parametrize("tag", ["a2_small", "a2_rect", "a2_L1", "a2_L1_peaked"])
def test_a2_log_optimal_transport(tag):
    g = load_golden("ot")
    out = oracle.log_optimal_transport(g[tag + "_scores"], float(g[tag + "_alpha"]), g[tag + "_ns"], 100)
    assert out.shape == g[tag + "_out"].shape
    np.testing.assert_allclose(out, g[tag + "_out"], atol=OT_TOL, rtol=0)
    # match-index parity: row/col argmax identical (est_position, first_layer.py:162)
    assert (out.argmax(2) == g[tag + "_out"].argmax(2)).all()
    assert (out.argmax(1) == g[tag + "_out"].argmax(1)).all()


def test_a2_single_iteration():
    g = load_golden("ot")
    out = oracle.log_optimal_transport(g["a2_small_scores"], float(g["a2_small_alpha"]), g["a2_small_ns"], 1)
    np.testing.assert_allclose(out, g["a2_small_out_it1"], atol=OT_TOL, rtol=0)


@pytest.mark.parametrize("tag", ["a3_L2", "a3_L3", "a3_L3_wide", "a3_rect"])
def test_a3_log_optimal_transport2(tag):
    g = load_golden("ot")
    out = oracle.log_optimal_transport2(g[tag + "_scores"], 1.0, g[tag + "_ns"], 100)
    np.testing.assert_allclose(out, g[tag + "_out"], atol=OT_TOL, rtol=0)
    assert (out.argmax(2) == g[tag + "_out"].argmax(2)).all()


def test_a4_tensor_resize():
    g = load_golden("resize")
    out = oracle.tensor_resize(g["src"].astype(np.float32), g["bound"])
    # values are 0..255; ATen's vectorised CPU lerp may differ in the last ulp
    np.testing.assert_allclose(out, g["out"], atol=6e-5, rtol=0)
    import torch

    gen = torch.Generator().manual_seed(int(g["src2_seed"]))
    src2 = torch.floor(torch.rand(1, 3, 736, 896, generator=gen) * 256).numpy()
    out2 = oracle.tensor_resize(src2, g["bound2"])
    np.testing.assert_allclose(out2, g["out2"], atol=6e-5, rtol=0)


def test_a4_rejects_bad_crop():
    src = np.zeros((1, 3, 8, 8), np.float32)
    with pytest.raises(RuntimeError):
        oracle.tensor_resize(src, np.array([[4, 4, 0, 3, 0]]))  # zero rows (library.cpp narrow would throw)
    with pytest.raises(RuntimeError):
        oracle.tensor_resize(src, np.array([[0, 4, 0, 8, 0]]))  # x1 past the edge


def test_a6_origin_extract():
    g = load_golden("subdivide")
    for tag in ("u8", "f32"):
        B, h, w, ps = g[f"ext_{tag}_dims"]
        out = oracle.origin_extract(g[f"ext_{tag}_left"], int(ps), int(w), int(h))
        assert out.dtype == g[f"ext_{tag}_out"].dtype
        assert np.array_equal(out, g[f"ext_{tag}_out"])


def test_a5_compute_imgs():
    g = load_golden("subdivide")
    h, w = (int(v) for v in g["ci_hw"])
    nl, nr, xs, ys, avg, bound5 = oracle.compute_imgs(g["ci_x_scale"], g["ci_y_scale"], g["ci_avg"], g["ci_nm"], g["ci_left"],
                                                      g["ci_right"], width=w, height=h)
    assert np.array_equal(nl, g["ci_new_left"])
    np.testing.assert_allclose(nr, g["ci_new_right"], atol=6e-5, rtol=0)
    assert np.array_equal(xs, g["ci_x_scale_new"])
    assert np.array_equal(ys, g["ci_y_scale_new"])
    assert np.array_equal(avg, g["ci_average_new"])


def test_a7_split_patches():
    g = load_golden("subdivide")
    for k in range(int(g["sp_count"])):
        hh, ww, mx = (int(v) for v in g[f"sp{k}_args"])
        cn, s2, s3 = oracle.split_patches(g[f"sp{k}_sum_cycle"], hh, ww, mx)
        assert cn == int(g[f"sp{k}_cycle_num"])
        assert np.array_equal(np.array(s2), g[f"sp{k}_second"])
        assert np.array_equal(np.array(s3), g[f"sp{k}_third"])


@pytest.mark.parametrize("tag,gh,gw,lb,it", [("L1", 15, 20, 1e-5, 15), ("L2", 12, 12, 1e-3, 8)])
def test_a8_iterative_expand_matrix(tag, gh, gw, lb, it):
    g = load_golden("expand")
    Z = g[tag + "_Z"]
    whole, core, avg, xs, ys, bound, nm = oracle.iterative_expand_matrix(np.exp(Z), g[tag + "_scalex"], g[tag + "_scaley"], gh, gw,
                                                                         lower_bound=lb, iter_num=it)
    assert np.array_equal(bound, g[tag + "_bound"])
    assert (bound[..., 1] > bound[..., 0]).any() and (bound[..., 3] > bound[..., 2]).any(), "fixture must grow boxes"
    np.testing.assert_allclose(avg, g[tag + "_average_point"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(xs, g[tag + "_x_scale"], rtol=2e-5)
    np.testing.assert_allclose(ys, g[tag + "_y_scale"], rtol=2e-5)
    np.testing.assert_allclose(whole, g[tag + "_whole_cost"].reshape(whole.shape), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(core, g[tag + "_core_cost"].reshape(core.shape), rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("tag,dust", [("L1", 300), ("L2", 144)])
def test_a9_est_nomatching(tag, dust):
    g = load_golden("expand")
    nm1, nm2 = oracle.est_nomatching(g[tag + "_Z"], dust)
    assert np.array_equal(nm1, g[tag + "_nm1"])
    assert np.array_equal(nm2, g[tag + "_nm2"])


@pytest.mark.parametrize("tag,merge_new", [("new", True), ("old", False)])
def test_a11_merge_patches(tag, merge_new):
    g = load_golden("merge")
    h, w = (int(v) for v in g["hw"])
    for c in (0, 1):
        out, sb, trust_after, nm2_after = oracle.merge_patches(merge_new, g[f"{tag}{c}_trust"], [32 * h, 32 * w], g[f"{tag}{c}_nm1"],
                                                               g[f"{tag}{c}_nm2"], g[f"{tag}{c}_sb_in"])
        assert np.array_equal(trust_after, g[f"{tag}{c}_trust_after"])
        assert np.array_equal(nm2_after, g[f"{tag}{c}_nm2_after"])
        assert np.array_equal(sb, g[f"{tag}{c}_sb_out"])
        assert np.array_equal(out, g[f"{tag}{c}_out"]), (tag, c, int((out != g[f"{tag}{c}_out"]).sum()))
        assert (~out).sum() > 0


def test_a14_get_result():
    g = load_golden("result")
    P = g["nm1"].shape[0]
    sc1 = np.repeat(g["sc1_first"][:, None, :], 2304, axis=1)
    ml, mr = oracle.get_result([g["nm0"], g["nm1"]], [g["pt0"], g["pt1"].astype(np.float32)], [g["sc0"], sc1],
                               [[32, 15, 20], [2, 48, 48]])
    assert ml.shape == g["matches_l"].shape and P > 0
    assert np.array_equal(ml, g["matches_l"])
    assert np.array_equal(mr, g["matches_r"])


def test_a13_third_compute_result():
    g = load_golden("third")
    scale = g["scale"]
    sx = np.sqrt(scale + np.float32(1e-8))
    m0, m1, im = oracle.third_compute_result(np.exp(g["Z"]), sx, sx, g["p_s"], g["p_t"])
    assert np.array_equal(im, g["if_matching1"])
    assert (~im).any() and im.any()
    assert np.array_equal(m0, g["mkpts0_f"])
    np.testing.assert_allclose(m1, g["mkpts1_f"], rtol=1e-5, atol=2e-5)


def test_a10_grid_sample12():
    from conftest import gathers_inputs

    g, maps, _, _ = gathers_inputs()
    out = oracle.grid_sample12(maps[0].numpy(), maps[1].numpy(), maps[2].numpy())
    assert np.array_equal(out, g["gs_out"])


def test_a12_third_unfold():
    from conftest import gathers_inputs

    g, _, feat0, feat1 = gathers_inputs()
    kenc = g["un_kenc"].reshape(128, 64)
    o0 = oracle.third_unfold(feat0.numpy(), g["un_mk0"], g["un_b"], kenc, g["un_rubbish"], g["un_mk0"], False)
    o1 = oracle.third_unfold(feat1.numpy(), g["un_mk1"], g["un_b"], kenc, g["un_rubbish"], g["un_mk0"], True)
    assert np.array_equal(o0, g["un_out0"]) and np.array_equal(o1, g["un_out1"])
    with pytest.raises(IndexError):  # a window whose flat index leaves the tensor: the reference's gather raises
        oracle.third_unfold(feat1.numpy(), np.array([[0.0, 0.0]], np.float32), np.array([0.0], np.float32), kenc, g["un_rubbish"],
                            g["un_mk0"][:1], True)
