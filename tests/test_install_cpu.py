"""install() rebinds the hot-path names inside the reference's caller modules (needs /root/reference; CPU only,
no compute is executed -- the replaced functions are CUDA-only)."""
import os
import sys

import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present (GPU box)")


def test_install_rebinds_callers_and_uninstall_restores():
    ref = ref_loader.load_reference()
    from pats_b200 import install as inst
    from pats_b200 import layers, modules, utils

    orig = ref.first_layer.log_optimal_transport
    orig_merge = ref.second_layer.SecondLayer.merge_patches_new
    done = inst.install()
    try:
        assert ("models.first_layer", "log_optimal_transport") in done
        assert ref.first_layer.log_optimal_transport is modules.log_optimal_transport
        assert ref.second_layer.log_optimal_transport2 is modules.log_optimal_transport2
        assert ref.third_layer.log_optimal_transport2 is modules.log_optimal_transport2
        assert ref.first_layer.Compute_imgs is utils.Compute_imgs
        assert ref.first_layer.Iterative_expand_matrix is utils.Iterative_expand_matrix
        assert ref.second_layer.Iterative_expand_matrix is utils.Iterative_expand_matrix
        assert ref.pats.get_result is utils.get_result
        assert ref.utils.tensor_resize.tensor_resize.__module__ == "pats_b200.tensor_resize"
        assert ref.second_layer.SecondLayer.merge_patches_new is layers.merge_patches_new
        assert ref.third_layer.ThirdLayer.Compute_result is layers.Compute_result
    finally:
        inst.uninstall()
    assert ref.first_layer.log_optimal_transport is orig
    assert ref.second_layer.SecondLayer.merge_patches_new is orig_merge


def test_install_fused_rebinds_the_layer_forwards_with_the_reference_signatures():
    """install(fused=True) also swaps SecondLayer.forward / ThirdLayer.forward for pats_b200.forward's mirrors: same parameter
    names in the same order as the reference's methods (models/second_layer.py:61, models/third_layer.py:112), restored by
    uninstall()."""
    import inspect

    ref = ref_loader.load_reference()
    from pats_b200 import forward as fwd
    from pats_b200 import install as inst

    o2, o3 = ref.second_layer.SecondLayer.forward, ref.third_layer.ThirdLayer.forward
    assert list(inspect.signature(fwd.second_layer_forward).parameters) == list(inspect.signature(o2).parameters)
    assert list(inspect.signature(fwd.third_layer_forward).parameters) == list(inspect.signature(o3).parameters)
    done = inst.install(fused=True)
    try:
        assert ("models.second_layer", "SecondLayer.forward") in done and ("models.third_layer", "ThirdLayer.forward") in done
        assert ref.second_layer.SecondLayer.forward is fwd.second_layer_forward
        assert ref.third_layer.ThirdLayer.forward is fwd.third_layer_forward
    finally:
        inst.uninstall()
    assert ref.second_layer.SecondLayer.forward is o2 and ref.third_layer.ThirdLayer.forward is o3
    plain = inst.install()
    try:
        assert ("models.second_layer", "SecondLayer.forward") not in plain
        assert ref.second_layer.SecondLayer.forward is o2
    finally:
        inst.uninstall()


def test_install_attention_rebinds_the_network_forward_and_refuses_the_cpu():
    """install(attention=True) swaps AttentionalGNN.forward (models/modules.py:126-134) for pats_b200.gnn's mirror -- same parameters,
    visible through every caller module's star import -- and uninstall() restores it; on CPU tensors the mirror raises (no fallback)."""
    import inspect

    import pytest
    import torch

    ref = ref_loader.load_reference()
    from pats_b200 import gnn as G
    from pats_b200 import install as inst

    orig = ref.modules.AttentionalGNN.forward
    assert list(inspect.signature(G.attentional_gnn_forward).parameters) == list(inspect.signature(orig).parameters)
    done = inst.install(attention=True)
    try:
        assert ("models.modules", "AttentionalGNN.forward") in done
        assert ref.second_layer.AttentionalGNN.forward is G.attentional_gnn_forward  # the class object the layers construct
        gnn = ref.modules.AttentionalGNN(16, ["self", "cross"]).eval()
        with pytest.raises(RuntimeError, match="CUDA-only"):
            gnn(torch.zeros(1, 16, 5), torch.zeros(1, 16, 5))
    finally:
        inst.uninstall()
    assert ref.modules.AttentionalGNN.forward is orig
    assert ("models.modules", "AttentionalGNN.forward") not in inst.install()
    inst.uninstall()
    assert G.supported(145, 264, 4) and G.supported(65, 128, 4) and G.supported(300, 448, 4) and G.supported(1024, 448, 4)
    assert not G.supported(65, 130, 4) and not G.supported(65, 1024, 4)
