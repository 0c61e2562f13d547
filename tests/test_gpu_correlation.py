"""The tcgen05 descriptor correlation (csrc/correlation.cu) against the reference's formulation: torch.einsum('bdn,bdm->bnm') in FP32
(allow_tf32 off, as torch defaults), then / sqrt(d) and * 0.1 (first_layer.py:110-114, second_layer.py:100-104,
third_layer.py:156-158).  3xTF32 must be FP32-accurate: the plans downstream are held to 1e-4."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(d0, d1):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s = torch.einsum('bdn,bdm->bnm', d0, d1)
        s = s / d0.shape[1] ** .5
        return 0.1 * s
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("b,d,n,m", [(37, 128, 65, 65), (300, 128, 65, 65), (9, 264, 145, 145), (2, 448, 300, 300), (3, 8, 8, 16), (2, 40, 130, 17),
                                     (1, 448, 480, 400)])
def test_correlation_matches_fp32_einsum(b, d, n, m):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from pats_b200 import layers as Ly

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(b * 1000 + d + n)
    d0 = (torch.randn(b, d, n, generator=g) * 3.0).to(dev)
    d1 = (torch.randn(b, d, m, generator=g) * 3.0).to(dev)
    out = Ly.correlation(d0, d1, 0.1 / math.sqrt(d))
    torch.cuda.synchronize()
    ref = _ref(d0, d1)
    ref64 = 0.1 * torch.einsum('bdn,bdm->bnm', d0.double(), d1.double()) / math.sqrt(d)
    err = float((out.double() - ref64).abs().max())
    err_ref = float((ref.double() - ref64).abs().max())
    assert out.shape == ref.shape
    # FP32-class accuracy: within a small factor of the FP32 GEMM's own error (measured 1 - 5x: the tensor core's accumulator does not
    # round to nearest, which shows over the 168 accumulation steps of d = 448) and below 6e-6 of the largest score -- two orders
    # of magnitude inside what the 1e-4 plan tolerance needs (second test)
    assert err <= max(8.0 * err_ref, 2e-6), f"max |ours - f64| = {err:.3e}, FP32 einsum: {err_ref:.3e}"
    assert err <= 6e-6 * float(ref64.abs().max()), f"relative error {err / float(ref64.abs().max()):.2e}"
    assert float((out - ref).abs().max()) < 4e-5


def test_correlation_feeds_the_sinkhorn_with_identical_matches():
    """level-3 shaped: Z from the tcgen05 kernel vs Z from the reference formulation -> plans within 1e-4, same row / column argmax."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from pats_b200 import layers as Ly, modules as M

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    b = 256
    d0 = torch.randn(b, 128, 65, generator=g).to(dev) * 2.0
    d1 = torch.randn(b, 128, 65, generator=g).to(dev) * 2.0
    ns = torch.exp((torch.rand(b, 1, 64, generator=g) * 2 - 1) * 1.0).to(dev)
    one = torch.tensor(1.0, device=dev)
    Za = M.log_optimal_transport2(Ly.correlation(d0, d1, 0.1 / math.sqrt(128)), one, ns, 100)
    Zb = M.log_optimal_transport2(_ref(d0, d1), one, ns, 100)
    assert float((Za - Zb).abs().max()) <= 1e-4
    assert torch.equal(Za.argmax(2), Zb.argmax(2)) and torch.equal(Za.argmax(1), Zb.argmax(1))
