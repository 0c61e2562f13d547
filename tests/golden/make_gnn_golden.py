"""Generate tests/golden/gnn.npz by running the UNMODIFIED reference `models.modules.AttentionalGNN` (this container only).

    python tests/golden/make_gnn_golden.py

The weights come from oracle.gnn.seeded_params (a numpy RNG: the fixture stores the seed, not 45 MB of weights) and are loaded
into the reference module through its own `load_state_dict`; inputs are seeded; outputs are the reference's, on CPU in float32 and
in float64 (`.double()` on the same module).  Cases: a tiny one, the level-3 shape (D = 128, n = 65, 10 layers,
third_layer.py:90), the level-2 shape (D = 264, n = 145, 18 layers, second_layer.py:42-43) and the level-1 shape (D = 448, n = 300,
first_layer.py).  Reference entry points: models/modules.py:119 AttentionalGNN, :108 AttentionalPropagation, :90
MultiHeadedAttention, :84 attention, :58 MLP.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
from ref_loader import load_reference  # noqa: E402
from oracle import gnn as O  # noqa: E402

CASES = {  # name: (seed, B, D, N, layer names)
    "tiny": (11, 2, 16, 7, ["self", "cross"] * 2),
    "l3": (12, 3, 128, 65, ["self", "cross"] * 5),
    "l2": (13, 2, 264, 145, ["self", "cross"] * 9),
    "l1": (14, 1, 448, 300, ["self", "cross"] * 9),
}


def inputs(seed, B, D, N):
    rng = np.random.default_rng(seed + 1000)
    return rng.standard_normal((B, D, N)).astype(np.float32), rng.standard_normal((B, D, N)).astype(np.float32)


def main():
    ref = load_reference()
    out = {}
    for name, (seed, B, D, N, names) in CASES.items():
        params = O.seeded_params(seed, len(names), D)
        gnn = ref.modules.AttentionalGNN(D, names).eval()
        sd = {}
        for l, p in enumerate(params):
            for k, v in p.items():
                sd[f"layers.{l}.{k}"] = torch.from_numpy(v)
            sd[f"layers.{l}.mlp.1.num_batches_tracked"] = torch.tensor(0)
        gnn.load_state_dict(sd, strict=True)
        d0, d1 = inputs(seed, B, D, N)
        with torch.no_grad():
            o0, o1 = gnn(torch.from_numpy(d0), torch.from_numpy(d1))
            g64 = gnn.double()
            p0, p1 = g64(torch.from_numpy(d0).double(), torch.from_numpy(d1).double())
        out[name + "_meta"] = np.array([seed, B, D, N, len(names)], dtype=np.int64)
        out[name + "_out0_f32"], out[name + "_out1_f32"] = o0.numpy(), o1.numpy()
        if name in ("tiny", "l3"):  # the larger cases keep the float32 outputs only (fixture size); f32 vs f64 is printed below
            out[name + "_out0_f64"], out[name + "_out1_f64"] = p0.numpy(), p1.numpy()
        print(name, "max|out|", float(p0.abs().max()), "f32 vs f64", float((o0.double() - p0).abs().max()))
    path = os.path.join(HERE, "gnn.npz")
    np.savez_compressed(path, **out)
    print(f"gnn.npz: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
