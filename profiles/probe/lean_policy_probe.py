"""Policy rollout without outputs (config 3): k_rollout (123 registers, 7 blocks per SM) against k_rollout_lean
(72 registers, 14 blocks per SM).  Run once per library: SUPPLYNET_LIB=build/libsupplynet_nolean.so for the former."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from multi_echelon_drl_b200.register_envs import make
from multi_echelon_drl_b200.simulator import BasePolicyParams
pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
out = {"lib": os.path.basename(os.environ.get("SUPPLYNET_LIB", "libsupplynet.so"))}
def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)
for n in (65536, 131072, 262144, 1048576):
    env = make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=3)
    sim = env.sim
    env._next_episode()
    sim.track_stats = True
    sim.stats.zero_()
    sim.episode(policy=pol, write_state=False)
    torch.cuda.synchronize()
    out[f"stats_{n}"] = [float(x) for x in sim.stats.cpu()[:4]]
    out[f"ms_{n}"] = timed(lambda: sim.episode(policy=pol, write_state=False))
    if n == 131072:
        out["loop_graph_100_per_episode"] = timed(lambda: env.run_policy_episodes(pol, 100, use_graph=True), 3) / 100
    del env, sim
print(json.dumps(out))
