"""N>1 path on CPU: two gloo ranks shard a batch by env index (no data-path collective) and
all-reduce the statistics vector.  The per-shard statistics are produced by the oracle here (the
checker standing in for the kernels, which need a GPU); the test pins the host-side logic: the shard
partition, the reduction, and shard invariance (2 ranks == 1 rank)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402

N, T = 64, 100


def _inputs():
    rs = np.random.RandomState(21)
    init = np.concatenate([rs.randint(0, 25, (N, 4)), rs.randint(0, 9, (N, 15))], axis=1)
    demand = rs.randint(0, 9, (N, T + 3))
    acts = rs.randint(0, 17, (T, N, 4))
    return init, demand, acts


def _shard_stats(lo, hi):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_rollout
    from multi_echelon_drl_b200.distributed import add_episode_moments
    init, demand, acts = _inputs()
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init[lo:hi], demand[lo:hi], acts[:, lo:hi])
    st = torch.zeros(16, dtype=torch.float64)
    st[0] = (hi - lo) * T
    st[1] = ref["rew"].sum()
    st[2:6] = torch.as_tensor(ref["holding"].sum((0, 1)))
    st[6:10] = torch.as_tensor(ref["backorder"].sum((0, 1)))
    add_episode_moments(st, torch.as_tensor(ref["rew"].sum(0)))
    return st


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from multi_echelon_drl_b200.distributed import allreduce_stats, shard_range
    lo, hi = shard_range(N, rank, world)
    st = allreduce_stats(_shard_stats(lo, hi))
    if rank == 0:
        torch.save(st, out)
    dist.destroy_process_group()


def test_two_rank_shards_reduce_to_single_rank_stats(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    want = _shard_stats(0, N)
    assert torch.equal(got, want)            # integer-valued sums: exact in float64, order-independent
    from multi_echelon_drl_b200.distributed import summarize
    s = summarize(got)
    assert s["env_steps"] == N * T and s["episodes"] == N
    assert abs(s["mean_episode_return"] - want[11].item() / N) < 1e-9


def test_allreduce_stats_is_identity_without_process_group():
    from multi_echelon_drl_b200.distributed import allreduce_stats
    st = torch.arange(16, dtype=torch.float64)
    assert torch.equal(allreduce_stats(st.clone()), st)
    with pytest.raises(ValueError):
        allreduce_stats(torch.zeros(3, dtype=torch.float64))


def _peer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from multi_echelon_drl_b200.distributed import PeerStatsAllReduce, allreduce_stats
    try:
        PeerStatsAllReduce(torch.device("cpu"))
        msg = "constructed"
    except RuntimeError as exc:
        msg = str(exc)
    # the group is still usable afterwards: the NCCL / gloo form carries the vector
    st = allreduce_stats(torch.full((16,), float(rank + 1), dtype=torch.float64))
    torch.save((msg, st), f"{out}.{rank}")
    dist.destroy_process_group()


def test_peer_memory_allreduce_refuses_together_where_there_is_no_device(tmp_path):
    """Without a CUDA device the peer-memory exchange cannot be built: every rank learns that in the same two
    collectives and raises (nobody is left waiting), and the process group keeps working for `allreduce_stats`
    (the fallback bench.py takes)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "peer")
    mp.spawn(_peer_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        msg, st = torch.load(f"{out}.{rank}")
        assert "cannot be mapped on every rank" in msg
        assert torch.equal(st, torch.full((16,), 3.0, dtype=torch.float64))
