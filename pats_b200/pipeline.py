"""The evaluation loop around `PATS.forward` with its CPU tail overlapped.

Mirrors the three loops of the reference's evaluate.py:20-39 / :42-61 / :65-84:

    for data in DataLoader(dataset, batch_size=1, shuffle=False, num_workers=0):     # cv2.imread + resize + pad in __getitem__
        data['image0'] = data['image0'].cuda(); data['image1'] = data['image1'].cuda()
        result = model(data)
        error_R, error_t = compute_pose_error(result['matches_l'].cpu().numpy(), result['matches_r'].cpu().numpy(),
                                              data['K0'][0].numpy(), data['K1'][0].numpy(), data['T0'][0].numpy(), data['T1'][0].numpy(),
                                              scale_factor, threshold)                # cv2.findEssentialMat (RANSAC) + recoverPose

The reference runs the three stages of a pair one after the other on one thread: the GPU idles while OpenCV decodes the next
images and while RANSAC works on the last matches.  With the hot path at ~1 ms and the forward pass at ~0.15 s these two CPU
stages are what bounds pairs/s next.  `evaluate_pairs` keeps the stages and their order, and runs them as a three-stage software
pipeline:

    loader thread    dataset[i] (its own code, unmodified) -> default_collate -> pinned host memory, `prefetch` pairs ahead
    caller's thread  images host -> device (non-blocking), model(data), match lists device -> pinned host (non-blocking) + event
    metrics thread   waits for the event, then pose_fn(...) -- ONE thread, pairs in order: OpenCV's RANSAC draws from a thread-local
                     generator whose state carries from call to call, so the sequence of calls on one fresh thread reproduces the
                     sequential loop's numbers exactly (a pool would not)

OpenCV and numpy release the GIL inside their kernels, so plain threads overlap; nothing is pickled.  Results come back in pair
order; an exception in any stage is re-raised in the caller.  Multi-GPU: give each rank its shard (`pats_b200.dist.shard_range`)
through `indices`.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch

__all__ = ["evaluate_pairs", "evaluate_pairs_sequential"]

_STOP = object()


def _collate(item):
    """What DataLoader(batch_size=1) hands the loop: a leading batch dimension, numpy -> tensor."""
    from torch.utils.data import default_collate

    return default_collate([item])


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.pin_memory() if torch.cuda.is_available() and not t.is_pinned() else t


def evaluate_pairs_sequential(model, dataset, pose_fn: Callable, scale_factor: float, threshold: float, device="cuda",
                              indices: Optional[Iterable[int]] = None) -> Tuple[List[float], List[float]]:
    """The reference's loop, stage after stage (evaluate.py:20-39) -- the baseline `evaluate_pairs` is measured against."""
    error_R_list, error_t_list = [], []
    for i in (range(len(dataset)) if indices is None else indices):
        data = _collate(dataset[i])
        data['image0'] = data['image0'].to(device)
        data['image1'] = data['image1'].to(device)
        result = model(data)
        error_R, error_t = pose_fn(result['matches_l'].cpu().numpy(), result['matches_r'].cpu().numpy(), data['K0'][0].numpy(), data['K1'][0].numpy(),
                                   data['T0'][0].numpy(), data['T1'][0].numpy(), scale_factor, threshold)
        error_R_list.append(error_R)
        error_t_list.append(error_t)
    return error_R_list, error_t_list


def evaluate_pairs(model, dataset, pose_fn: Callable, scale_factor: float, threshold: float, device="cuda",
                   indices: Optional[Sequence[int]] = None, prefetch: int = 2, stats: Optional[dict] = None) -> Tuple[List[float], List[float]]:
    """Same inputs, same calls, same order of results as `evaluate_pairs_sequential`; loading and pose estimation overlap the
    forward pass.  `stats` (optional dict) receives per-stage busy seconds."""
    idx = list(range(len(dataset)) if indices is None else indices)
    n = len(idx)
    on_gpu = torch.device(device).type == "cuda"
    loaded: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))
    to_pose: "queue.Queue" = queue.Queue()
    errors: List[Optional[BaseException]] = [None, None]
    results: List[Optional[tuple]] = [None] * n
    busy = {"load_s": 0.0, "pose_s": 0.0}
    stop = threading.Event()

    def put(q, item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def loader():
        import time

        try:
            for i in idx:
                t0 = time.perf_counter()
                data = _collate(dataset[i])
                if on_gpu:
                    data['image0'], data['image1'] = _pin(data['image0']), _pin(data['image1'])
                busy["load_s"] += time.perf_counter() - t0
                if not put(loaded, data):
                    return
        except BaseException as e:  # noqa: BLE001  (re-raised in the caller)
            errors[0] = e
        finally:
            put(loaded, _STOP)

    def metrics():
        import time

        try:
            while True:
                job = to_pose.get()
                if job is _STOP:
                    return
                k, ml, mr, ev, data = job
                if ev is not None:
                    ev.synchronize()
                t0 = time.perf_counter()
                results[k] = pose_fn(ml.numpy(), mr.numpy(), data['K0'][0].numpy(), data['K1'][0].numpy(), data['T0'][0].numpy(), data['T1'][0].numpy(),
                                     scale_factor, threshold)
                busy["pose_s"] += time.perf_counter() - t0
        except BaseException as e:  # noqa: BLE001
            errors[1] = e
            stop.set()

    def next_loaded():
        # never blocks for good: a stage that failed sets `stop` (or leaves its error) and may not be able to queue its end marker
        while True:
            try:
                return loaded.get(timeout=0.1)
            except queue.Empty:
                if stop.is_set() or errors[0] is not None or not threads[0].is_alive():
                    try:
                        return loaded.get_nowait()
                    except queue.Empty:
                        return _STOP

    threads = [threading.Thread(target=loader, name="pats-loader", daemon=True), threading.Thread(target=metrics, name="pats-metrics", daemon=True)]
    for t in threads:
        t.start()
    try:
        for k in range(n):
            data = next_loaded()
            if data is _STOP or stop.is_set():
                break
            data['image0'] = data['image0'].to(device, non_blocking=True)
            data['image1'] = data['image1'].to(device, non_blocking=True)
            result = model(data)
            ml, mr = result['matches_l'], result['matches_r']
            if on_gpu and ml.is_cuda:
                hl = torch.empty(ml.shape, dtype=ml.dtype, pin_memory=True).copy_(ml, non_blocking=True)
                hr = torch.empty(mr.shape, dtype=mr.dtype, pin_memory=True).copy_(mr, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(ml.device))
            else:
                hl, hr, ev = ml.cpu(), mr.cpu(), None
            to_pose.put((k, hl, hr, ev, data))
    finally:
        stop_now = errors[0] is not None or errors[1] is not None
        if stop_now:
            stop.set()
        to_pose.put(_STOP)
        threads[1].join()
        stop.set()  # releases a loader blocked on a full queue
        while True:  # drain so the loader's final put succeeds
            try:
                loaded.get_nowait()
            except queue.Empty:
                break
        threads[0].join(timeout=5.0)
    for e in errors:
        if e is not None:
            raise e
    if any(r is None for r in results):
        raise RuntimeError("evaluate_pairs: the pipeline stopped before every pair was processed")
    if stats is not None:
        stats.update(busy)
    return [r[0] for r in results], [r[1] for r in results]
