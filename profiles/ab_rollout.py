"""A/B micro-benchmark of the rollout kernel (config-2 shape) for a given libsupplynet build.
Usage: SUPPLYNET_LIB=path/to/lib.so python profiles/ab_rollout.py [n_envs]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs, T
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork, BasePolicyParams
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dtype = sys.argv[2] if len(sys.argv) > 2 else "int32"
spec, init, demand, actions = make_inputs(n, 1)
sim = BatchedSupplyNetwork(spec, n, "cuda", dtype)
sim.load_episode(init, demand)
a = torch.as_tensor(actions).cuda().to(sim.tdtype)
to = torch.empty((T, n, 28), dtype=torch.float32, device="cuda"); tr = torch.empty((T, n), dtype=torch.float32, device="cuda")
pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        sim.reset(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2]
def f1(): sim.period, sim.terminal = 0, False; sim.rollout(T, actions=a, traj_obs=to, traj_reward=tr)
def f2(): sim.period, sim.terminal = 0, False; sim.rollout(T, policy=pol)
def f3(): sim.period, sim.terminal = 0, False; sim.rollout(T, actions=a)
def f4(): sim.period, sim.terminal = 0, False; sim.rollout(T, actions=a, traj_obs=to)
def f5(): sim.period, sim.terminal = 0, False; sim.rollout(T, actions=a, traj_reward=tr)
m1, m2, m3, m4, m5 = timeit(f1), timeit(f2), timeit(f3), timeit(f4), timeit(f5)
print(json.dumps({"lib": os.environ.get("SUPPLYNET_LIB", "default"), "n": n, "dtype": dtype, "traj_ms": m1, "traj_GBs": n*T*137/m1/1e6,
                  "heur_ms": m2, "heur_Gsteps": n*T/m2/1e6, "actions_only_ms": m3, "obs_only_ms": m4, "reward_only_ms": m5}))
