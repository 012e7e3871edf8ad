"""Multi-GPU plumbing: the env batch is sharded across ranks (one process per GPU); the step has NO collective.
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in CPU tests) is used only to all-gather per-rank
episode statistics at report time (BASELINE.json north_star)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous global env-id range [lo, hi) owned by ``rank``; earlier ranks take the remainder."""
    base, rem = divmod(int(num_envs_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_episode_stats(returns: torch.Tensor, lengths: torch.Tensor) -> Dict[str, float]:
    """All-gather (sum, sum of squares, count, length sum) of finished-episode returns across ranks and reduce them to
    global statistics.  ``returns``/``lengths``: 1-D tensors of this rank's finished episodes (may be empty)."""
    r = returns.double()
    return gather_episode_sums(float(r.sum()), float((r * r).sum()), float(r.numel()), float(lengths.double().sum()),
                               r.device)


def gather_episode_sums(s: float, ss: float, n: float, length_sum: float, device) -> Dict[str, float]:
    """The same from this rank's four sums (what the step kernels accumulate per env): the job's only collective."""
    local = torch.tensor([s, ss, n, length_sum], dtype=torch.float64, device=device)
    rank, size = world()
    if size > 1:
        parts = [torch.zeros_like(local) for _ in range(size)]
        dist.all_gather(parts, local)
        total = torch.stack(parts).sum(0)
    else:
        total = local
    s, ss, n, ln = (float(v) for v in total)
    mean = s / n if n else float("nan")
    var = max(ss / n - mean * mean, 0.0) if n else float("nan")
    return {"episodes": n, "return_mean": mean, "return_std": var ** 0.5 if n else float("nan"),
            "length_mean": ln / n if n else float("nan"), "world_size": size}


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device: int = 0, sysfs: str = "/sys/bus/pci/devices") -> Dict[str, object]:
    """Pin the calling process to the CPUs that are local to ``cuda:device`` (its PCIe root's NUMA node), so that pinned
    host buffers allocated afterwards are NUMA-local to the GPU that reads and writes them.  On a multi-socket 8-GPU box
    the host-buffer path (``wf_step_host``) otherwise pushes half of its traffic across the socket interconnect.
    Best effort: returns what was done; never raises."""
    import os

    info: Dict[str, object] = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as fp:
            cpus = _parse_cpulist(fp.read())
        try:
            with open(os.path.join(sysfs, bdf, "numa_node")) as fp:
                info["numa_node"] = int(fp.read().strip())
        except OSError:
            pass
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), pci=bdf)
    except Exception as exc:  # noqa: BLE001 - best effort by design
        info["error"] = repr(exc)
    return info
