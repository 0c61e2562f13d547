"""Multi-GPU plumbing: image pairs are independent (evaluate.py:25-35 loops pair by pair), so ranks take
contiguous blocks of the pair list and the only exchange is one gather of the per-pair match lists at the
end.  torch.distributed; NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/world) items per rank, last ranks may be short/empty."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_items // world)
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_match_lists(matches: Sequence[torch.Tensor], group=None, max_pairs: int | None = None, dst: int | None = None,
                       stats: dict | None = None) -> List[List[torch.Tensor]]:
    """Gather per-pair match lists (the path's ONE exchange: after the pair loop).

    matches: this rank's list of [K_i, D] float tensors (one per local pair; D = 4 for (yl,xl,yr,xr)), any K_i >= 0.
    max_pairs: an upper bound on the pairs per rank that every rank knows without talking (the shard size,
        `shard_range`); None costs one extra all_reduce to find it.
    dst: None -> every rank receives everything (all_gather); r -> only rank r does (what evaluate.py needs: metrics are
        computed in one place), the others get [].
    Returns a list over ranks of lists of tensors in pair order.

    Exactly two collectives and ONE host synchronisation, whatever the world size: a fixed-size int64 header per rank
    [n_pairs, K_0 .. K_{max_pairs-1}] (all_gather_into_tensor, read back with a single D2H copy), then one payload padded to
    the largest rank total.  Nothing here runs inside the pair loop.  `stats` (optional dict) receives payload_bytes, rows.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    dev = matches[0].device if len(matches) else (torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu"))
    D = matches[0].shape[1] if len(matches) else 4
    if max_pairs is None:
        n_max = torch.tensor([len(matches)], dtype=torch.int64, device=dev)
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=group)
        max_pairs = int(n_max.item())
    if len(matches) > max_pairs:
        raise ValueError(f"{len(matches)} local pairs exceed max_pairs = {max_pairs}")
    header = torch.zeros(1 + max_pairs, dtype=torch.int64)
    header[0] = len(matches)
    for i, m in enumerate(matches):  # shapes are host integers: no device round trip
        header[1 + i] = m.shape[0]
    header = header.to(dev)
    headers = torch.empty(world * (1 + max_pairs), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(headers, header, group=group)
    heads = headers.cpu().reshape(world, 1 + max_pairs)  # the one host synchronisation
    totals = heads[:, 1:].sum(1)
    max_rows = max(1, int(totals.max()))
    payload = torch.empty((max_rows, D), dtype=torch.float32, device=dev)
    if len(matches):
        torch.cat([m.to(torch.float32) for m in matches], 0, out=payload[: int(totals[rank])])
    if stats is not None:
        stats.update(payload_bytes_per_rank=max_rows * D * 4, rows_local=int(totals[rank]), rows_total=int(totals.sum()), collectives=2, host_syncs=1)
    if dst is None:
        recv = torch.empty((world * max_rows, D), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(recv, payload, group=group)
        parts = recv.reshape(world, max_rows, D)
    else:
        parts = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
        dist.gather(payload, parts, dst=dst, group=group)
        if rank != dst:
            return []
    out = []
    for r in range(world):
        n = int(heads[r, 0])
        rows, off = [], 0
        for c in heads[r, 1:1 + n].tolist():
            rows.append(parts[r][off:off + c])
            off += c
        out.append(rows)
    return out


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs `local_cpulist` of its PCI function).

    One process per GPU stages its inputs in pinned host memory; pinned pages are placed on the node of the allocating
    thread, so a rank that runs on the far socket pulls every host->device byte over the socket interconnect and all ranks
    together exhaust one socket's memory channels.  Call this BEFORE allocating pinned buffers.  Returns what was done
    ({"bound": False, "why": ...} when the topology cannot be read); never raises.
    """
    import os

    info = {"bound": False}
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(base + "/local_cpulist") as f:
            local = _parse_cpulist(f.read())
        node = -1
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except OSError:
            pass
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        info.update({"pci": bdf, "numa_node": node, "cpus_local": len(local), "cpus_allowed": len(allowed)})
        if not target:
            info["why"] = "no allowed CPU on the GPU's node"
            return info
        if target == allowed:
            info["why"] = "already local (single node or pre-bound)"
            return info
        os.sched_setaffinity(0, target)
        info["bound"] = True
        info["cpus"] = len(target)
    except Exception as e:  # no sysfs entry, no permission, not Linux ...
        info["why"] = f"{type(e).__name__}: {e}"
    return info
