"""Synthetic two-view geometry for the evaluation-loop tests and bench.py's `tail` leg (TEST / BASELINE INFRASTRUCTURE): a dataset
whose items have the keys and dtypes the reference's datasets return (datasets/megadepth.py, yfcc.py, scannet.py: `image0`, `image1`
uint8 [H,W,3]; `K0`, `K1` float32 [3,3]; `T0`, `T1` float32 [4,4]) plus planted correspondences of a random rigid scene, and a
stand-in for `PATS.forward` that returns those correspondences as `matches_l` / `matches_r` ((y, x) rows, as utils/utils.py:189-213
produces them) with a configurable share of outliers."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def reference_metrics():
    """The reference's utils/metrics.py (compute_pose_error :21-66, aggregate_metrics :89-95), unmodified."""
    import ref_loader

    if not ref_loader.reference_available():
        return None
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    return importlib.import_module("utils.metrics")


def _rot(rng, max_deg):
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    a = np.deg2rad(rng.uniform(-max_deg, max_deg))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx


class SyntheticTwoView(torch.utils.data.Dataset):
    def __init__(self, n_pairs=6, hw=(480, 640), n_points=400, outliers=0.3, seed=7, load_cost=0):
        self.n, self.hw, self.np_, self.out, self.seed, self.load_cost = n_pairs, hw, n_points, outliers, seed, load_cost

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed + i)
        H, W = self.hw
        # the images of the live-forward tests (tests/live_util.synthetic_pair: uniform noise, second image rolled by (16, 24)), so that
        # a real network in the GPU stage finds the matches it finds there
        g = torch.Generator().manual_seed(18027 + self.seed + i)
        img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8).numpy()
        scratch = img
        for _ in range(self.load_cost):  # stands in for cv2.imread + resize of the real datasets (CPU work that releases the GIL)
            import cv2

            scratch = cv2.GaussianBlur(scratch, (5, 5), 1.0)
        f = 0.9 * W
        K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float64)
        R, t = _rot(rng, 12.0), rng.uniform(-0.4, 0.4, 3)
        T1 = np.eye(4)
        T1[:3, :3], T1[:3, 3] = R, t
        X = np.stack([rng.uniform(-2, 2, self.np_), rng.uniform(-1.5, 1.5, self.np_), rng.uniform(4, 9, self.np_)], 1)
        x0 = (K @ X.T).T
        x0 = x0[:, :2] / x0[:, 2:]
        X1 = (R @ X.T).T + t
        x1 = (K @ X1.T).T
        x1 = x1[:, :2] / x1[:, 2:]
        bad = rng.random(self.np_) < self.out
        x1[bad] = np.stack([rng.uniform(0, W, bad.sum()), rng.uniform(0, H, bad.sum())], 1)
        return {"image0": img, "image1": np.roll(img, (16, 24), (0, 1)).copy(), "K0": K.astype(np.float32), "K1": K.astype(np.float32),
                "T0": np.eye(4, dtype=np.float32), "T1": T1.astype(np.float32),
                "planted_l": x0[:, ::-1].astype(np.float32).copy(), "planted_r": x1[:, ::-1].astype(np.float32).copy()}  # (y, x)


class PlantedModel:
    """`model(data)` -> {'matches_l', 'matches_r'} on the device of the images; optionally runs a real model first (its output is
    discarded: random-init weights give matches no essential matrix fits) so that the GPU stage has its real cost."""

    def __init__(self, real=None):
        self.real = real

    def __call__(self, data):
        dev = data["image0"].device
        if self.real is not None:
            r = self.real({"image0": data["image0"], "image1": data["image1"]})
            self.real_matches = getattr(self, "real_matches", 0) + int(r["matches_l"].shape[0])
        return {"matches_l": data["planted_l"][0].to(dev), "matches_r": data["planted_r"][0].to(dev)}
