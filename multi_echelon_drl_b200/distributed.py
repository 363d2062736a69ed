"""Multi-GPU plumbing: envs are independent, so they are sharded by index over ranks (one process
per GPU) with NO data-path collective.  The only exchange is an all-reduce (sum) of the small
statistics vector the kernels accumulate (episode return / per-facility cost totals) - NCCL over
NVLink on GPUs, gloo in the CPU tests.

Statistics vector layout (SN_STATS_WORDS = 16 doubles, include/supplynet.h):
  [0] env-steps  [1] sum reward  [2..5] sum holding cost per facility  [6..9] sum backorder cost per
  facility  [10] episodes  [11] sum episode return  [12] sum episode return^2  [13..15] reserved
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

STATS_WORDS = 16


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of env indices owned by `rank`: [lo, hi).  Remainders go to the low ranks so
    that sizes differ by at most one and the concatenation over ranks is the identity order."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def add_episode_moments(stats: torch.Tensor, ep_return: torch.Tensor) -> torch.Tensor:
    """Fold finished episodes' returns into words 10..12 (count, sum, sum of squares)."""
    r = ep_return.to(torch.float64)
    stats[10] += r.numel()
    stats[11] += r.sum()
    stats[12] += (r * r).sum()
    return stats


def allreduce_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the statistics vector over ranks (in place).  No-op without an initialised process group."""
    if stats.numel() != STATS_WORDS or stats.dtype != torch.float64:
        raise ValueError("stats must be a float64 vector of SN_STATS_WORDS entries")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().cpu().tolist()
    steps, eps = max(s[0], 1.0), max(s[10], 1.0)
    mean = s[11] / eps
    return {"env_steps": s[0], "mean_reward": s[1] / steps, "holding_cost": s[2:6], "backorder_cost": s[6:10],
            "episodes": s[10], "mean_episode_return": mean,
            "std_episode_return": max(s[12] / eps - mean * mean, 0.0) ** 0.5}


def bind_to_gpu_cpus(device_index: int) -> Optional[list]:
    """Pin this process to the CPUs next to GPU `device_index` (NVML's ideal CPU affinity = the socket the GPU's PCIe
    root hangs off), so that pinned host buffers allocated afterwards live in that socket's memory and the H2D / D2H
    streams of the host-buffer path do not cross the inter-socket link.  One process per GPU calls this before it
    allocates pinned memory.  Returns the CPU list, or None if nothing was changed (NVML missing, cpuset too narrow)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device_index)      # CUDA index -> PCI address -> NVML handle
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08X}:{p.pci_bus_id:02X}:{p.pci_device_id:02X}.0".encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        ideal = {c for c in range(n_cpu) if (int(words[c // 64]) >> (c % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(ideal & allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
