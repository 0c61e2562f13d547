"""The attention network's kernels (csrc/gnn.cu) for compute-sanitizer: every GEMM variant (TMA-fed in CTA pairs with multicast, TMA-fed
single CTAs, register-staged), both precisions, the resident-key attention kernels of both generations, the chunked-key (flash)
kernel, and the train()-mode BatchNorm path -- on tiny batches (the tools are 10-100x slower than a plain run).

    compute-sanitizer --tool memcheck  python tools/sanitize_gnn.py
    compute-sanitizer --tool synccheck python tools/sanitize_gnn.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import gnn as O  # noqa: E402
from pats_b200 import _lib, gnn as G  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
order = ("attn.proj.0", "attn.proj.1", "attn.proj.2", "attn.merge", "mlp.0")
g = torch.Generator().manual_seed(5)
for (L_, D_, N_, B_) in ((2, 128, 65, 3), (2, 264, 145, 2), (2, 448, 300, 1)):
    params = O.seeded_params(9, L_, D_)
    raw = np.concatenate([np.concatenate([np.concatenate([p[k + ".weight"].reshape(-1), p[k + ".bias"]]) for k in order]
                                         + [p["mlp.1.weight"], p["mlp.1.bias"], p["mlp.1.running_mean"], p["mlp.1.running_var"],
                                            p["mlp.3.weight"].reshape(-1), p["mlp.3.bias"]]) for p in params])
    raw_d = torch.from_numpy(raw).to(dev)
    packed = G.pack_raw(raw_d, L_, D_, 4, O.BN_EPS)
    x0, x1 = torch.randn(B_, D_, N_, generator=g).to(dev), torch.randn(B_, D_, N_, generator=g).to(dev)
    cross = bytes([0, 1])
    ref = None
    for gv in (0, 2, 1):
        lib.pats_gnn_gemm_variant(gv)
        for av in (0, 1):
            lib.pats_gnn_attention_variant(av)
            for passes in (3, 1):
                G.set_precision(passes)
                o0, _ = G.attentional_gnn(packed, cross, 4, x0, x1)
                if passes == 3:
                    ref = o0 if ref is None else ref
                    assert torch.equal(ref, o0), (D_, gv, av)
    lib.pats_gnn_gemm_variant(0), lib.pats_gnn_attention_variant(0), G.set_precision(3)
    # train()-mode BatchNorm
    packed_t = torch.empty(lib.pats_gnn_packed_floats(L_, D_), dtype=torch.float32, device=dev)
    _lib.check(lib.pats_gnn_pack_train_f32(raw_d.data_ptr(), L_, D_, 4, packed_t.data_ptr(), None), "pack_train")
    running = torch.zeros(L_, 2, 2 * D_, device=dev)
    running[:, 1] = 1.0
    ws = torch.empty(lib.pats_gnn_workspace_floats(B_, D_, N_), device=dev)
    out0, out1 = torch.empty_like(x0), torch.empty_like(x1)
    _lib.check(lib.pats_attentional_gnn_train_f32(x0.data_ptr(), x1.data_ptr(), B_, D_, N_, packed_t.data_ptr(), raw_d.data_ptr(), running.data_ptr(), 0.1, 1e-5,
                                                  cross, L_, 4, out0.data_ptr(), out1.data_ptr(), ws.data_ptr(), ws.numel(), None), "train")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out0).all())
    print(f"D={D_} n={N_}: ok", flush=True)
print("sanitize_gnn: done")
