"""Which index does torch.argsort(x)[..., 0] (unstable, CUDA) return among tied minima of a 9-wide slice?

second_layer.py:169,230 picks the owner window of every fine cell with `torch.argsort(scores)[:, :, :, 0]` over nine candidates;
ties at the minimum are the rule, not the exception (absent windows score exactly 0.0; matched cells are `-10000 + trust` in f32,
quantised to ~1e-3).  On CUDA tensors ATen sorts slices of <= 32 elements with an unstable bitonic network
(ATen/native/cuda/SortUtils.cuh: bitonicSortKVInPlace, 32 slots, 16 threads, invalid slots sort to the end), so the winner
among equal keys is decided by the network's exchange pattern.  This probe runs the emulation of that network (the same
function the merge kernel's table was generated from) against the live op on random tie patterns.

    python tools/argsort_tie_probe.py            # on a GPU box; prints the agreement and writes gpurun_out/argsort_tie_probe.json
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bitonic_first(keys, n_valid=9, size32=32, lt=True):
    """Emulates bitonicSort<32> on rows of `keys` [N, n_valid]; returns the original index that ends in slot 0."""
    N = keys.shape[0]
    k = np.full((N, size32), np.inf)
    k[:, :n_valid] = keys
    v = np.tile(np.arange(size32), (N, 1))
    ok = np.zeros((N, size32), bool)
    ok[:, :n_valid] = True
    T = size32 // 2

    def step(stride, flags):
        for t in range(T):
            pos = 2 * t - (t & (stride - 1))
            a, b = pos, pos + stride
            with np.errstate(invalid="ignore"):
                c = (k[:, a] < k[:, b]) if lt else (k[:, a] > k[:, b])
            swap = (c & ok[:, a]) | ~ok[:, b]
            do = swap == flags[t]
            for arr in (k, v, ok):
                ta = arr[do, a].copy()
                arr[do, a] = arr[do, b]
                arr[do, b] = ta

    size = 2
    while size < size32:
        flags = [(t & (size // 2)) != 0 for t in range(T)]
        stride = size // 2
        while stride > 0:
            step(stride, flags)
            stride //= 2
        size *= 2
    stride = size32 // 2
    while stride > 0:
        step(stride, [False] * T)
        stride //= 2
    return v[:, 0]


def main():
    import torch

    dev = "cuda:0"
    rng = np.random.default_rng(0)
    N = 200000
    pool = np.array([0.0, 0.0, 0.0, 1e-14, 8e-14, 100000.0, -9998.912109375, -9998.5, 0.25, 1.5])
    x = pool[rng.integers(0, len(pool), size=(N, 9))]
    x[: N // 4] = 0.0  # all-equal rows
    got = torch.argsort(torch.from_numpy(x).to(dev).reshape(1, 400, 500, 9))[..., 0].reshape(-1).cpu().numpy()
    res = {}
    for lt in (True, False):
        emu = bitonic_first(x, lt=lt)
        res["LTOp" if lt else "GTOp"] = float((emu == got).mean())
    first = np.argmin(x, 1)
    res["first_index"] = float((first == got).mean())
    res["rows"] = N
    res["torch"] = torch.__version__
    # the winner depends only on WHICH of the nine slots hold the minimum: tabulate it for all 511 non-empty masks
    table = []
    for m in range(1, 512):
        row = np.where([(m >> i) & 1 for i in range(9)], 0.0, 1.0)[None]
        table.append(int(bitonic_first(row, lt=True)[0]))
    xm = (x == x.min(1, keepdims=True))
    masks = (xm * (1 << np.arange(9))).sum(1)
    res["mask_table_agrees"] = float((np.array([0] + table)[masks] == got).mean())
    print(json.dumps(res))
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(REPO, "gpurun_out", "argsort_tie_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    if "table" in sys.argv:
        t = [0] + [int(bitonic_first(np.where([(m >> i) & 1 for i in range(9)], 0.0, 1.0)[None], lt=True)[0]) for m in range(1, 512)]
        print(t)
    else:
        main()
