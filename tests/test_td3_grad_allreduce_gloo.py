"""HgeTD3._allreduce_grads on two gloo ranks (CPU): every rank ends with the mean of the ranks' gradients."""
import os
import sys

import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import torch.nn as nn
    sys.path.insert(0, ROOT)
    from multi_echelon_drl_b200.td3 import HgeTD3
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 2))
    for i, p in enumerate(m.parameters()):
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1) + torch.arange(p.numel(), dtype=p.dtype).view_as(p) * rank
    HgeTD3._allreduce_grads(m)
    ok = True
    for i, p in enumerate(m.parameters()):
        want = sum(torch.full_like(p, float(r + 1)) * (i + 1) + torch.arange(p.numel(), dtype=p.dtype).view_as(p) * r
                   for r in range(world)) / world
        ok = ok and torch.allclose(p.grad, want)
    out[rank] = ok
    dist.destroy_process_group()


def test_gradients_are_averaged_over_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
