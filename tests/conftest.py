import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def gathers_inputs():
    """Regenerates the large seeded inputs of tests/golden/gathers.npz exactly as make_golden.gen_gathers drew them."""
    import torch

    g = load_golden("gathers")
    gF = torch.Generator().manual_seed(int(g["gs_seed"]))
    N, P = int(g["gs_N"]), int(g["un_P"])
    maps = [torch.randn(N, 64, 48, 48, generator=gF), torch.randn(N, 64, 24, 24, generator=gF), torch.randn(N, 128, 12, 12, generator=gF)]
    maps = [m.half().float() for m in maps]
    feat0 = torch.randn(P, 128, 52, 52, generator=gF).half().float()
    feat1 = torch.randn(P, 128, 52, 52, generator=gF).half().float()
    return g, maps, feat0, feat1
