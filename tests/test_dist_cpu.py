"""world_size-2 gloo test of the multi-GPU plumbing (pair sharding + gather of match lists)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pats_b200.dist import gather_match_lists, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 1500):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, mode="all"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_pairs, rank, world)
        mine = []
        for i in range(lo, hi):
            g = torch.Generator().manual_seed(i)
            k = int(torch.randint(0, 50, (1,), generator=g))
            mine.append(torch.rand(k, 4, generator=g))
        stats = {}
        if mode == "all":
            allm = gather_match_lists(mine, stats=stats)
        elif mode == "bounded":  # the shard size is known to every rank: two collectives, no extra reduction
            allm = gather_match_lists(mine, max_pairs=-(-n_pairs // world), stats=stats)
        else:  # gather to rank 0 only (evaluate.py computes its metrics in one place)
            allm = gather_match_lists(mine, max_pairs=-(-n_pairs // world), dst=0, stats=stats)
            if rank != 0:
                assert allm == []
                return
        assert stats["collectives"] == 2 and stats["host_syncs"] == 1
        assert len(allm) == world
        flat = [m for per_rank in allm for m in per_rank]
        assert len(flat) == n_pairs
        for i, m in enumerate(flat):
            g = torch.Generator().manual_seed(i)
            k = int(torch.randint(0, 50, (1,), generator=g))
            assert m.shape == (k, 4)
            assert torch.equal(m, torch.rand(k, 4, generator=g))
    finally:
        dist.destroy_process_group()


def test_gather_match_lists_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port(), 5), nprocs=2, join=True)


def test_gather_match_lists_uneven_and_empty_rank():
    mp.spawn(_worker, args=(2, _free_port(), 1), nprocs=2, join=True)


def test_gather_match_lists_bounded_header_and_gather_to_rank0():
    mp.spawn(_worker, args=(2, _free_port(), 7, "bounded"), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, _free_port(), 7, "dst"), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, _free_port(), 0, "bounded"), nprocs=2, join=True)


def test_parse_cpulist_and_numa_binding_never_raises():
    """bind_host_to_gpu reads the GPU's sysfs `local_cpulist`; without a GPU (here) it must report, not raise."""
    from pats_b200.dist import _parse_cpulist, bind_host_to_gpu

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    info = bind_host_to_gpu(0)
    assert info["bound"] is False and "why" in info


def _pipeline_worker(rank, world, port, n_pairs, out_dir):
    """Each rank runs its shard of the pair list through the overlapped evaluation loop (pats_b200.pipeline, N4); the pose errors are
    gathered with all_gather_object and compared, on rank 0, with ONE sequential loop over its own shard (the RANSAC stream is per
    shard: a rank starts a fresh metrics thread)."""
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pose_util as P
    from pats_b200 import pipeline as PL

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = P.reference_metrics()
        ds = P.SyntheticTwoView(n_pairs=n_pairs, n_points=200, outliers=0.3)
        lo, hi = shard_range(n_pairs, rank, world)
        mine = PL.evaluate_pairs(P.PlantedModel(), ds, m.compute_pose_error, 1.0, 0.5, device="cpu", indices=list(range(lo, hi)))
        everyone = [None] * world
        dist.all_gather_object(everyone, (lo, hi, mine))
        if rank == 0:
            assert [e[0] for e in everyone] == [shard_range(n_pairs, r, world)[0] for r in range(world)]
            assert sum(len(e[2][0]) for e in everyone) == n_pairs
            import threading

            box = {}
            t = threading.Thread(target=lambda: box.setdefault("r", PL.evaluate_pairs_sequential(P.PlantedModel(), ds, m.compute_pose_error, 1.0, 0.5,
                                                                                              device="cpu", indices=range(lo, hi))))
            t.start()
            t.join()
            assert np.array_equal(np.array(box["r"][0]), np.array(mine[0])) and np.array_equal(np.array(box["r"][1]), np.array(mine[1]))
            errs = [x for e in everyone for x in e[2][0]]
            assert all(np.isfinite(errs)) and max(errs) < 5.0
            open(os.path.join(out_dir, "ok"), "w").write("1")
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_loop_gloo_world2(tmp_path):
    import pytest

    import pose_util as P

    if P.reference_metrics() is None:
        pytest.skip("reference Python not available")
    mp.spawn(_pipeline_worker, args=(2, _free_port(), 5, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
