"""The numpy restatement of the reference's attention network (oracle/gnn.py, models/modules.py:58-134) against
tests/golden/gnn.npz: outputs of the unmodified `models.modules.AttentionalGNN` on seeded weights and inputs
(tests/golden/make_gnn_golden.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
GOLD = os.path.join(REPO, "tests", "golden", "gnn.npz")


def case(g, name):
    from make_gnn_golden import CASES, inputs
    from oracle import gnn as O

    seed, B, D, N, names = CASES[name]
    assert list(g[name + "_meta"]) == [seed, B, D, N, len(names)]
    return O.seeded_params(seed, len(names), D), names, inputs(seed, B, D, N)


@pytest.mark.parametrize("name", ["tiny", "l3", "l2"])
def test_oracle_matches_reference_outputs(name):
    from oracle import gnn as O

    g = np.load(GOLD)
    params, names, (d0, d1) = case(g, name)
    o0, o1 = O.attentional_gnn(params, names, d0, d1)
    scale = max(np.abs(g[name + "_out0_f32"]).max(), np.abs(g[name + "_out1_f32"]).max())
    # float64 restatement vs the reference's float32 run: the reference's own rounding (2e-6 at these depths, printed by the generator)
    assert np.abs(o0 - g[name + "_out0_f32"]).max() <= 2e-5 * scale
    assert np.abs(o1 - g[name + "_out1_f32"]).max() <= 2e-5 * scale
    if name + "_out0_f64" in g.files:  # the reference module run in float64: restatement error only
        assert np.abs(o0 - g[name + "_out0_f64"]).max() <= 1e-11 * scale
        assert np.abs(o1 - g[name + "_out1_f64"]).max() <= 1e-11 * scale
