"""Multi-GPU plumbing: envs are independent, so they are sharded by index over ranks (one process
per GPU) with NO data-path collective.  The only exchange is an all-reduce (sum) of the small
statistics vector the kernels accumulate (episode return / per-facility cost totals): on GPUs one
kernel over peer memory (`PeerStatsAllReduce`: sn_stats_allreduce, mailboxes mapped over NVLink) or
NCCL (`allreduce_stats`), gloo in the CPU tests.

Statistics vector layout (SN_STATS_WORDS = 16 doubles, include/supplynet.h):
  [0] env-steps  [1] sum reward  [2..5] sum holding cost per facility  [6..9] sum backorder cost per
  facility  [10] episodes  [11] sum episode return  [12] sum episode return^2  [13..15] reserved
"""
from __future__ import annotations

import contextlib
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


class PeerStatsAllReduce:
    """All-reduce (sum) of the statistics vector as ONE kernel launch on the caller's stream (include/supplynet.h,
    `sn_stats_allreduce`): every rank stores its 16 doubles straight into every peer's mailbox over NVLink (entries that
    carry value and sequence tag together), polls its own mailbox until all ranks' entries of this call are there and
    sums them in rank order - bit-identical on every rank.  No copy, no second stream, no host synchronisation; next
    to a 150 us episode kernel it costs a few microseconds where the NCCL call (copy kernel + event hand-over +
    collective kernel that does not fit beside the rollout's blocks) costs 25-30.

    One instance per process and process group; every rank makes the same calls in the same order on one stream.
    Construction exchanges the mailboxes' IPC handles through the process group (all_gather_object) and raises if a
    peer's memory cannot be mapped (callers then keep `allreduce_stats`)."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import ctypes as C
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerStatsAllReduce needs an initialised process group")
        self._lib, self._C, self._check = _lib.load(), C, _lib.check
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise NotImplementedError("PeerStatsAllReduce: at most 16 ranks (one node)")
        self._peers, self._mail, self.seq = [], None, 0
        cuda = self.device.type == "cuda" and torch.cuda.is_available()
        with (torch.cuda.device(self.device) if cuda else contextlib.nullcontext()):
            # a rank that fails early still takes part in both exchanges below: nobody is left waiting in a collective
            mail, handle, err = C.c_void_p(), (C.c_ubyte * 64)(), None
            if not cuda:
                err = "no CUDA device"
            elif self._lib.sn_mail_alloc(self.world, C.byref(mail), handle) != 0:
                err = self._lib.sn_last_error().decode()
            else:
                self._mail = mail
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r, h in enumerate(handles):
                if err is not None:
                    break
                if r == self.rank:
                    ptrs.append(mail.value)
                    continue
                peer = C.c_void_p()
                rc = self._lib.sn_mail_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(peer)) if any(h) else -1
                if rc != 0:
                    err = err or (self._lib.sn_last_error().decode() if any(h) else "the peer could not export its mailbox")
                    continue
                self._peers.append(peer)
                ptrs.append(peer.value)
            # everybody learns whether everybody succeeded: a half-built exchange must not be used by anyone
            ok = [None] * self.world
            dist.all_gather_object(ok, err is None, group=group)
            if not all(ok):
                self.close(barrier=False)
                raise RuntimeError(f"PeerStatsAllReduce: peer mailboxes cannot be mapped on every rank ({err or 'another rank failed'})")
            self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def __call__(self, stats: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sum `stats` (float64 [16], this device) over the ranks into `out` (default: in place), on the current stream."""
        out = stats if out is None else out
        for t in (stats, out):
            if t.numel() != STATS_WORDS or t.dtype != torch.float64 or t.device != self.device or not t.is_contiguous():
                raise ValueError("stats must be contiguous float64 vectors of SN_STATS_WORDS entries on this rank's device")
        self.seq = self.seq % 0xFFFFFFFF + 1                        # 1, 2, ..., never 0 (the tag of an empty mailbox)
        C = self._C
        self._check(self._lib.sn_stats_allreduce(
            C.c_void_p(stats.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(self._ptrs.data_ptr()), self.rank,
            self.world, self.seq, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def close(self, barrier: bool = True):
        """Unmaps the peers' mailboxes and frees the own one (after a barrier: nobody may still be writing to it)."""
        if self._mail is None and not self._peers:
            return
        torch.cuda.synchronize(self.device)
        if barrier:
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._peers:
                self._lib.sn_mail_close(p)
            self._peers = []
            if barrier:
                dist.barrier(group=self.group)                      # every peer has unmapped it
            if self._mail is not None:
                self._lib.sn_mail_free(self._mail)
        self._mail = None


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
