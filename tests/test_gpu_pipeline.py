"""The overlapped evaluation loop (pats_b200/pipeline.py) on the GPU: the unmodified `PATS.forward` with the whole
path installed as the GPU stage, the reference's `compute_pose_error` (utils/metrics.py:21-66, OpenCV RANSAC) as the CPU tail; pose
errors identical to the reference's sequential loop (evaluate.py:20-39), bit for bit."""
import threading

import numpy as np
import pytest
import torch

import live_util as L
import pose_util as P

pytestmark = pytest.mark.gpu


def test_pipeline_on_the_gpu_equals_the_sequential_loop():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    m = P.reference_metrics()
    if m is None or L.reference_root() is None:
        pytest.skip("reference Python not staged")
    import pats_b200.install as inst
    from pats_b200 import pipeline as PL

    dev = torch.device("cuda:0")
    ref = L.load_reference()
    ds = P.SyntheticTwoView(n_pairs=4, n_points=500, outliers=0.3, load_cost=2)
    with torch.no_grad():
        real = L.build_model(ref, L.config(if_local=True, merge_new=True), device=dev)
        inst.install(fused=True, attention=True)
        try:
            model = P.PlantedModel(real=real)
            box = {}
            t = threading.Thread(target=lambda: box.setdefault("r", PL.evaluate_pairs_sequential(model, ds, m.compute_pose_error, 1.0, 0.5, device=dev)))
            t.start()
            t.join()
            stats = {}
            par = PL.evaluate_pairs(model, ds, m.compute_pose_error, 1.0, 0.5, device=dev, stats=stats)
        finally:
            inst.uninstall()
    seq = box["r"]
    assert np.array_equal(np.array(seq[0]), np.array(par[0])) and np.array_equal(np.array(seq[1]), np.array(par[1]))
    assert all(np.isfinite(par[0])) and max(par[0]) < 5.0
    assert stats["pose_s"] > 0
