"""Shard invariance across REAL devices ("N envs on 1 GPU == same N on 8 GPUs"): one global batch, cut by
env index over all visible GPUs, gives bit for bit what the whole batch gives on one GPU.  Needs at least two GPUs
(skipped on the single-GPU box of the round-end test run; `bench.py --gpus N` makes the same check under torchrun at
every N > 1 and reports it as `shard_invariance`)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_index_sharding_over_devices_equals_one_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two or more GPUs")
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.distributed import shard_range
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    G, n, T = torch.cuda.device_count(), 10007, 100
    spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
    rs = np.random.RandomState(5)
    init, demand = bulk_episode_inputs(spec, n, rs)
    acts = rs.randint(0, 17, (T, n, 4)).astype(np.int32)

    def run(dev, lo, hi):
        with torch.cuda.device(dev):
            sim = BatchedSupplyNetwork(spec, hi - lo, f"cuda:{dev}", "int32")
            sim.load_episode(init[lo:hi], demand[lo:hi])
            out = sim.episode(T, actions=torch.as_tensor(acts[:, lo:hi]), record_obs=True, record_reward=True)
            return out["traj_obs"].cpu(), out["traj_reward"].cpu(), sim.ep_return.cpu()

    whole = run(0, 0, n)
    parts = [run(g, *shard_range(n, g, G)) for g in range(G)]
    assert torch.equal(torch.cat([p[0] for p in parts], dim=1), whole[0])
    assert torch.equal(torch.cat([p[1] for p in parts], dim=1), whole[1])
    assert torch.equal(torch.cat([p[2] for p in parts]), whole[2])
