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


def gather_match_lists(matches: Sequence[torch.Tensor], group=None) -> List[List[torch.Tensor]]:
    """All-gather per-pair match lists.

    matches: this rank's list of [K_i, D] float tensors (one per local pair; D = 4 for (yl,xl,yr,xr)).
    Returns, on every rank, a list over ranks of lists of tensors in pair order.
    Two collectives: the per-pair counts, then one padded payload.
    """
    world = dist.get_world_size(group)
    dev = matches[0].device if len(matches) else torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    D = matches[0].shape[1] if len(matches) else 4
    n_local = torch.tensor([len(matches)], dtype=torch.int64, device=dev)
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(n_all, n_local, group=group)
    max_pairs = max(int(t.item()) for t in n_all)
    counts = torch.zeros(max(max_pairs, 1), dtype=torch.int64, device=dev)
    if len(matches):
        counts[: len(matches)] = torch.tensor([m.shape[0] for m in matches], dtype=torch.int64, device=dev)
    counts_all = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(counts_all, counts, group=group)
    max_rows = max(1, max(int(c.sum().item()) for c in counts_all))
    payload = torch.zeros((max_rows, D), dtype=torch.float32, device=dev)
    if len(matches):
        cat = torch.cat([m.to(torch.float32) for m in matches], 0)
        payload[: cat.shape[0]] = cat
    payload_all = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(payload_all, payload, group=group)
    out = []
    for r in range(world):
        n = int(n_all[r].item())
        cs = counts_all[r][:n].tolist()
        rows, off = [], 0
        for c in cs:
            rows.append(payload_all[r][off:off + c])
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
