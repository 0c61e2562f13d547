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
