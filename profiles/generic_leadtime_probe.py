import sys, json; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dataclasses import replace
from multi_echelon_drl_b200 import uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork, BasePolicyParams
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
n, T = 65536, 100
for li, ls in (((2,2,2,1),(2,2,2,2)), ((1,1,1,1),(1,1,1,1)), ((3,2,1,2),(2,3,4,1))):
    spec = replace(uniform_spec(None, True, T), info_leadtime=li, ship_leadtime=ls)
    sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
    init, demand = bulk_episode_inputs(spec, n, np.random.RandomState(0))
    sim.load_episode(init, demand)
    a = torch.randint(0, 17, (T, n, 4), device="cuda", dtype=torch.int32)
    to = torch.empty((T, n, 28), device="cuda"); tr = torch.empty((T, n), device="cuda")
    pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / 10
    m1 = t(lambda: sim.episode(T, actions=a, traj_obs=to, traj_reward=tr, write_state=False))
    m2 = t(lambda: sim.episode(T, policy=pol, write_state=False))
    print(json.dumps({"li": li, "ls": ls, "state_words": sim.state_words, "traj_ms": m1, "policy_ms": m2, "traj_GBs": n*T*137/m1/1e6}))
