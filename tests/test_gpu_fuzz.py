"""Randomised differential campaign: CUDA path vs the C form of the oracle on large batches with random
scenario / agent set / observability / rule and adversarial inputs (orders far outside the action space,
negative orders, empty and overfull pipelines, zero and huge demand).  Bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import DPA  # noqa: E402
from oracle import beergame_c as OC  # noqa: E402
from oracle import beergame_np as O  # noqa: E402

AGENT_SETS = [(0, 1, 2, 3), (3, 2, 1, 0), (0,), (1,), (2,), (3,), (0, 1), (1, 2), (2, 3), (0, 1, 2), (1, 2, 3), (2, 1)]


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_against_c_oracle(seed):
    from multi_echelon_drl_b200 import normal_spec, uniform_spec
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    rs = np.random.RandomState(1000 + seed)
    scenario = ("uniform", "normal")[rs.randint(2)]
    agents = AGENT_SETS[rs.randint(len(AGENT_SETS))]
    glob = bool(rs.randint(2))
    rule = ("a", "d+a")[rs.randint(2)]
    T = int(rs.choice([1, 2, 7, 100]))
    n = int(rs.choice([40000, 65536, 100001]))
    spec = (uniform_spec if scenario == "uniform" else normal_spec)(agents, glob, T)
    if rule == "d+a":
        spec = spec.with_rule("d+a", *DPA[scenario])
    ospec = O.scenario_spec(scenario, agents, glob, T, rule)
    # adversarial inputs: a quarter of the envs start empty, a quarter overfull; demand mixes 0, typical and spikes
    init = np.concatenate([rs.randint(0, 60, (n, 4)), rs.randint(0, 40, (n, 15))], axis=1)
    init[: n // 4] = 0
    init[n // 4: n // 2] *= 10
    demand = rs.randint(0, 12, (n, T + 3))
    demand[rs.rand(n, T + 3) < 0.05] = 0
    demand[rs.rand(n, T + 3) < 0.02] = 500
    acts = rs.randint(-5, 40, (T, n, len(agents)))
    acts[rs.rand(T, n, len(agents)) < 0.01] = 2000           # far outside the action space: never clipped in rule 'a'
    ref = OC.rollout(ospec, init, demand, acts, n_threads=8)
    sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
    sim.load_episode(init, demand)
    obs0 = torch.zeros((n, sim.obs_dim), device="cuda")
    out = sim.episode(T, actions=torch.as_tensor(acts), record_obs=True, record_reward=True, obs0_out=obs0)
    assert np.array_equal(obs0.cpu().numpy(), ref["obs0"]), (scenario, agents, glob, rule, T, n)
    assert np.array_equal(out["traj_obs"].cpu().numpy(), ref["obs"]), (scenario, agents, glob, rule, T, n)
    assert np.array_equal(sim.ep_return.cpu().numpy(), ref["ep_return"])
    assert np.array_equal(out["traj_reward"].cpu().numpy(), ref["rew"].astype(np.float32))
    # the single-period kernel on the same inputs (first min(T, 5) periods)
    sim.reset()
    r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(min(T, 5)):
        o, _, _ = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
        assert np.array_equal(o.cpu().numpy(), ref["obs"][t]) and np.array_equal(r64.cpu().numpy(), ref["rew"][t])
