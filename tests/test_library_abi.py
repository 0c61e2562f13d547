"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the torch-facing wrappers refuse to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO

HEADER = os.path.join(REPO, "include", "pats_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pats_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from pats_b200 import build

    so = build.build()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    from pats_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.pats_version() == 100


def test_kernel_dispatch_table():
    from pats_b200 import _lib

    lib = _lib.load()
    assert lib.pats_sinkhorn_kernel_kind(65, 65) == 0      # level 3: one warp per problem
    assert lib.pats_sinkhorn_kernel_kind(145, 145) == 1    # level 2: one CTA per problem
    assert lib.pats_sinkhorn_kernel_kind(301, 301) == 3    # level 1: one 8-CTA cluster per problem
    assert lib.pats_sinkhorn_kernel_kind(512, 512) == 3
    assert lib.pats_sinkhorn_kernel_kind(1537, 1537) == 4  # rows split over co-resident CTAs, plan streamed
    assert lib.pats_sinkhorn_kernel_kind(1025, 1025) == 4
    assert lib.pats_sinkhorn_kernel_kind(4097, 4097) == 4
    assert lib.pats_sinkhorn_kernel_kind(5000, 5000) == 2  # generic log-domain kernel


def test_argument_validation_needs_no_gpu():
    from pats_b200 import _lib

    lib = _lib.load()
    rc = lib.pats_log_optimal_transport2_f32(None, None, None, 4, 0, 65, 100, None, None)
    assert rc == -1 and b"bad sizes" in lib.pats_last_error() or b"null" in lib.pats_last_error()
    rc = lib.pats_tensor_resize_f32(None, 1, 3, 0, 10, None, 2, 96, 96, None, None, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_wrappers_fail_loudly_without_cuda():
    from pats_b200 import modules, tensor_resize, utils

    s = torch.zeros(1, 5, 5)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        modules.log_optimal_transport2(s, 1.0, torch.ones(1, 1, 4), 10)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        modules.log_optimal_transport(s, 1.0, torch.ones(1, 1, 5), 10)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        tensor_resize.tensor_resize(torch.zeros(1, 3, 8, 8), torch.zeros(1, 5, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        utils.origin_extract(torch.zeros(1, 3, 12, 12), 4, 1, 1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pats_b200/ may reference it."""
    pkg = os.path.join(REPO, "pats_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libpats_oracle" not in text, f


def test_attention_network_sizes_and_validation_need_no_gpu():
    """pats_gnn_*: the size functions follow the layouts include/pats_b200.h documents; bad arguments are refused before any CUDA call."""
    from pats_b200 import _lib

    lib = _lib.load()
    L, D = 18, 264
    DD = D * D
    assert lib.pats_gnn_raw_floats(L, D) == L * (4 * (DD + D) + 4 * DD + 2 * D + 8 * D + 2 * DD + D)
    assert lib.pats_gnn_packed_floats(L, D) == L * (9 * DD + 6 * D + 18 * DD)  # FP32 layers, then their TF32 halves
    assert lib.pats_gnn_workspace_floats(3, D, 145) == 28 * 3 * 145 * D + 32 * D
    rc = lib.pats_attentional_gnn_f32(None, None, 1, 264, 145, None, None, 18, 4, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.pats_last_error()
    rc = lib.pats_attentional_gnn_f32(None, None, -1, 264, 145, None, None, 18, 4, None, None, None, 0, None)
    assert rc == -1 and b"bad sizes" in lib.pats_last_error()
    assert lib.pats_attentional_gnn_f32(None, None, 0, 264, 145, None, None, 18, 4, None, None, None, 0, None) == 0  # empty batch
    rc = lib.pats_gnn_pack_f32(None, 18, 264, 5, 1e-5, None, None)
    assert rc == -1 and b"bad sizes" in lib.pats_last_error()
    rc = lib.pats_attentional_gnn_train_f32(None, None, 1, 264, 145, None, None, None, 0.1, 1e-5, None, 18, 4, None, None, None, 0, None)
    assert rc == -1
