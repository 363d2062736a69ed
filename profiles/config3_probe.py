"""Config-3 episode loop (131,072 envs, in-kernel base-stock policy): where does the time per episode go?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200.register_envs import make
from multi_echelon_drl_b200.simulator import BasePolicyParams
n = 131072
pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
out = {}
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)
env = make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=3)
sim = env.sim
env._next_episode()
for stats in (False, True):
    sim.track_stats = stats
    out[f"episode_kernel_stats_{stats}"] = timed(lambda: sim.episode(policy=pol, write_state=False))
sim.track_stats = False
out["fills_only"] = timed(lambda: env._device_episode())
out["loop_serial_per_episode"] = timed(lambda: env.run_policy_episodes(pol, 10, use_graph=False), 5) / 10
out["loop_graph_per_episode"] = timed(lambda: env.run_policy_episodes(pol, 10, use_graph=True), 5) / 10
out["loop_graph_100_per_episode"] = timed(lambda: env.run_policy_episodes(pol, 100, use_graph=True), 3) / 100
print(json.dumps(out))
