"""The library keeps its cached state per DEVICE (kernel attributes such as the 10-CTA non-portable cluster size and the dynamic
shared-memory limits, the fallback counter, the hand-over flag pools and the scratch buffers): the same process must be able to
use a second GPU after the first -- including the default stream, which is `0` on every device -- and get the same bits.
Needs two GPUs (`gpurun --gpus 2`); skipped otherwise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_second_device_in_one_process_gives_identical_results():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pats_b200 import _lib, layers as Ly, modules as M, utils as U

    lib = _lib.load()
    g = torch.Generator().manual_seed(12)
    s1 = 0.5 * torch.randn(1, 300, 300, generator=g)
    ns1 = torch.exp((torch.rand(1, 1, 300, generator=g) * 2 - 1) * 1.0)
    s2 = 0.5 * torch.randn(7, 145, 145, generator=g)
    sx, sy = torch.exp(torch.randn(7, 144, generator=g) * 0.3), torch.exp(torch.randn(7, 144, generator=g) * 0.3)
    s3 = 0.5 * torch.randn(33, 65, 65, generator=g)
    s3[5] *= 300.0  # a problem that takes the in-kernel fallback (touches the per-device counter)
    ns3 = torch.exp((torch.rand(33, 1, 64, generator=g) * 2 - 1) * 1.0)
    ps = torch.randint(0, 24, (33, 2), generator=g) * 4
    sb = 0.3 * torch.randn(2, 600, 600, generator=g)
    nsb = torch.exp((torch.rand(2, 1, 600, generator=g) * 2 - 1) * 1.0)
    nm1 = torch.rand(3, 2304, generator=g) < 0.8
    nm0 = torch.ones(1, 300, dtype=torch.bool)
    nm0[0, 4:7] = False
    avg = torch.rand(1, 300, 2, generator=g) * torch.tensor([13.0, 18.0]) + 1.0
    pt1 = torch.rand(3, 2304, 2, generator=g) * 48

    def run(dev):
        d = torch.device("cuda", dev)
        t = lambda x: x.to(d)  # noqa: E731
        out = []
        for stream in (None, torch.cuda.Stream(d)):  # the default stream (the same handle on every device) and a private one
            with torch.cuda.device(d), torch.cuda.stream(stream):
                one = torch.tensor(1.0, device=d)
                Z1 = M.log_optimal_transport(t(s1), one, t(ns1), 30)              # 10-CTA cluster kernel (non-portable size)
                e1 = Ly.est_position(Z1, t(ns1), t(ns1), 15, 20, 15, 1e-5)        # zeroed per-(device, stream) workspace
                m2 = Ly.second_layer_match(t(s2), 1.0, t((sx * sy).reshape(7, 1, 144)), t(sx), t(sy), 30, True, 12)  # hand-over flag pool
                sxy = (t(ns3).reshape(33, 64) + 1e-8).sqrt()
                m3 = Ly.third_layer_match(t(s3), 1.0, t(ns3), sxy, sxy, t(ps), t(ps), 30)
                Zb = M.log_optimal_transport(t(sb), one, t(nsb), 5)                # grid-cooperative kernel workspace
                ml, mr = U.get_result(1, [t(nm0), t(nm1)], [t(avg), t(pt1)], [torch.ones(1, 300, 2, device=d), torch.ones(3, 2304, 2, device=d)],
                                      [[32, 15, 20], [2, 48, 48]], None)
                torch.cuda.synchronize(d)
                fb = lib.pats_sinkhorn_fallback_count(1)
            out.append([x.cpu() for x in (Z1, *e1, *m2, *m3, Zb, ml, mr)] + [torch.tensor(fb)])
        return out

    a = run(0)
    b = run(1)
    again = run(0)
    for which, other in (("device 1", b), ("device 0 again", again)):
        for si in range(2):
            for i, (x, y) in enumerate(zip(a[si], other[si])):
                same = torch.equal(x, y) or (x.is_floating_point() and bool(((x == y) | (x.isnan() & y.isnan())).all()))
                assert same, f"{which}, stream {si}, output {i}: differs from the first run on device 0"
    assert int(a[0][-1]) >= 1, "the ill-conditioned problem did not reach the per-device fallback counter"
    assert np.isfinite(a[0][0].numpy()).all()


def test_attention_network_on_a_second_device():
    """csrc/gnn.cu: kernel attributes (dynamic shared memory of the GEMM / attention kernels) are set per device, tensor maps carry the
    device's pointers, cluster launches go to the current device -- the same network on device 1 after device 0 gives the same bits."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import gnn as O
    from pats_b200 import gnn as G

    L_, D_, N_, B_ = 4, 264, 145, 5
    params = O.seeded_params(3, L_, D_)
    order = ("attn.proj.0", "attn.proj.1", "attn.proj.2", "attn.merge", "mlp.0")
    raw = np.concatenate([np.concatenate([np.concatenate([p[k + ".weight"].reshape(-1), p[k + ".bias"]]) for k in order]
                                         + [p["mlp.1.weight"], p["mlp.1.bias"], p["mlp.1.running_mean"], p["mlp.1.running_var"],
                                            p["mlp.3.weight"].reshape(-1), p["mlp.3.bias"]]) for p in params])
    g = torch.Generator().manual_seed(4)
    x0, x1 = torch.randn(B_, D_, N_, generator=g), torch.randn(B_, D_, N_, generator=g)
    cross = bytes([0, 1] * (L_ // 2))

    def run(dev):
        d = torch.device("cuda", dev)
        with torch.cuda.device(d):
            packed = G.pack_raw(torch.from_numpy(raw).to(d), L_, D_, 4, O.BN_EPS)
            o0, o1 = G.attentional_gnn(packed, cross, 4, x0.to(d), x1.to(d))
            torch.cuda.synchronize(d)
        return o0.cpu(), o1.cpu()

    a, b, again = run(0), run(1), run(0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0], again[0]) and torch.equal(a[1], again[1])
    r0, _ = O.attentional_gnn(params, ["self", "cross"] * (L_ // 2), x0.numpy(), x1.numpy())
    assert float(np.abs(a[0].numpy() - r0).max()) <= 3e-5 * float(np.abs(r0).max())
