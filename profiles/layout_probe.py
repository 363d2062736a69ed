"""Rollout speed for the non-identity layouts (decentralized roles, local observation, reversed agents)."""
import sys, json, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_echelon_drl_b200 import uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
n, T = 65536, 100
for agents, glob in (((0, 1, 2, 3), True), ((3, 2, 1, 0), True), ((0,), False), ((0,), True), ((1, 2), False)):
    spec = uniform_spec(agents, glob, T)
    sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
    init, demand = bulk_episode_inputs(spec, n, np.random.RandomState(0))
    sim.load_episode(init, demand)
    a = torch.randint(0, 17, (T, n, len(agents)), device="cuda", dtype=torch.int32)
    to = torch.empty((T, n, sim.obs_dim), device="cuda"); tr = torch.empty((T, n), device="cuda")
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / 10
    m1 = t(lambda: sim.episode(T, actions=a, traj_obs=to, traj_reward=tr, write_state=False))
    m0 = t(lambda: sim.episode(T, actions=a, write_state=False))
    bytes_ = n * T * (4 * len(agents) + 4 + 4 * sim.obs_dim + 5)
    print(json.dumps({"agents": agents, "global": glob, "obs_dim": sim.obs_dim, "traj_ms": m1, "no_output_ms": m0,
                      "GBs": bytes_ / m1 / 1e6, "G_env_steps_s": n * T / m1 / 1e6}))
