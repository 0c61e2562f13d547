"""The overlapped evaluation loop (pats_b200/pipeline.py) against the reference's sequential loop
(evaluate.py:20-39) on the CPU: same pose errors, bit for bit and in order -- RANSAC's thread-local generator included -- with the
reference's own utils/metrics.py:21-66 as the pose function; exceptions of every stage surface in the caller."""
import numpy as np
import pytest
import torch

import pose_util as P


def _metrics():
    m = P.reference_metrics()
    if m is None:
        pytest.skip("reference Python not available")
    return m


def _sequential_in_fresh_thread(fn):
    """OpenCV's generator is per thread: the baseline runs on a fresh thread too, like the pipeline's metrics thread."""
    import threading

    box = {}
    t = threading.Thread(target=lambda: box.setdefault("r", fn()))
    t.start()
    t.join()
    return box["r"]


def test_pipeline_reproduces_the_sequential_loop_bit_for_bit():
    from pats_b200 import pipeline as PL

    m = _metrics()
    ds = P.SyntheticTwoView(n_pairs=7, n_points=300, outliers=0.35)
    model = P.PlantedModel()
    seq = _sequential_in_fresh_thread(lambda: PL.evaluate_pairs_sequential(model, ds, m.compute_pose_error, 1.0, 0.5, device="cpu"))
    stats = {}
    par = PL.evaluate_pairs(model, ds, m.compute_pose_error, 1.0, 0.5, device="cpu", prefetch=2, stats=stats)
    assert np.array_equal(np.array(seq[0]), np.array(par[0])) and np.array_equal(np.array(seq[1]), np.array(par[1]))
    assert all(np.isfinite(seq[0])) and max(seq[0]) < 5.0 and max(seq[1]) < 10.0  # the planted poses are recovered
    assert m.aggregate_metrics(*seq) == m.aggregate_metrics(*par)
    assert stats["pose_s"] > 0 and stats["load_s"] > 0
    # a shard of the pair list (multi-GPU: pats_b200.dist.shard_range)
    sub = PL.evaluate_pairs(model, ds, m.compute_pose_error, 1.0, 0.5, device="cpu", indices=[2, 3, 4])
    assert len(sub[0]) == 3


def test_pipeline_surfaces_errors_of_every_stage():
    from pats_b200 import pipeline as PL

    m = _metrics()
    ds = P.SyntheticTwoView(n_pairs=4, n_points=100)

    class BadDataset(P.SyntheticTwoView):
        def __getitem__(self, i):
            if i == 2:
                raise KeyError("broken pair")
            return super().__getitem__(i)

    with pytest.raises(KeyError):
        PL.evaluate_pairs(P.PlantedModel(), BadDataset(n_pairs=4, n_points=100), m.compute_pose_error, 1.0, 0.5, device="cpu")

    def bad_pose(*a):
        raise ValueError("pose")

    with pytest.raises(ValueError):
        PL.evaluate_pairs(P.PlantedModel(), ds, bad_pose, 1.0, 0.5, device="cpu")

    def bad_model(data):
        raise RuntimeError("forward")

    with pytest.raises(RuntimeError, match="forward"):
        PL.evaluate_pairs(bad_model, ds, m.compute_pose_error, 1.0, 0.5, device="cpu")


def test_too_few_matches_follow_the_reference():
    from pats_b200 import pipeline as PL

    m = _metrics()
    ds = P.SyntheticTwoView(n_pairs=2, n_points=10)  # < 15 matches: utils/metrics.py:23-24 returns (inf, inf)
    r = PL.evaluate_pairs(P.PlantedModel(), ds, m.compute_pose_error, 1.0, 0.5, device="cpu")
    assert r == ([np.inf, np.inf], [np.inf, np.inf])
