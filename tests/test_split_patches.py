"""a7 split_patches: host integer logic -- checked on the CPU against the reference's golden outputs and the oracle."""
import numpy as np
import torch

import oracle
from conftest import load_golden
from pats_b200.utils import split_patches


def test_split_patches_golden_and_oracle():
    g = load_golden("subdivide")
    for k in range(int(g["sp_count"])):
        hh, ww, mx = (int(v) for v in g[f"sp{k}_args"])
        sc = torch.from_numpy(g[f"sp{k}_sum_cycle"])
        cn, s2, s3 = split_patches(sc, hh, ww, mx)
        assert cn == int(g[f"sp{k}_cycle_num"])
        assert np.array_equal(np.array(s2), g[f"sp{k}_second"])
        assert np.array_equal(np.array(s3), g[f"sp{k}_third"])
    gen = torch.Generator().manual_seed(5)
    for _ in range(50):
        hh, ww = int(torch.randint(1, 40, (1,), generator=gen)), int(torch.randint(1, 40, (1,), generator=gen))
        frac = float(torch.rand(1, generator=gen))
        mx = int(torch.randint(1, 600, (1,), generator=gen))
        sc = torch.cumsum((torch.rand(hh * ww, generator=gen) < frac).int(), 0)
        assert split_patches(sc, hh, ww, mx) == oracle.split_patches(sc.numpy(), hh, ww, mx)
