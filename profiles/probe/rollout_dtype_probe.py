"""Config-2 rollout kernel (65,536 envs x 100, streamed actions, trajectory out) per quantity type, best of 5 x 20 launches."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from multi_echelon_drl_b200 import uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
n, T = 65536, 100
spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T)
out = {"lib": os.path.basename(os.environ.get("SUPPLYNET_LIB", "libsupplynet.so"))}
for dt in ("int32", "float32"):
    sim = BatchedSupplyNetwork(spec, n, "cuda", dt)
    sim.load_episode(*bulk_episode_inputs(spec, n, np.random.RandomState(0)))
    acts = torch.randint(0, 17, (T, n, 4), device="cuda", dtype=torch.int32).to(sim.tdtype)
    t_obs = torch.empty((T, n, 28), dtype=torch.float32, device="cuda"); t_rew = torch.empty((T, n), dtype=torch.float32, device="cuda")
    for _ in range(5): sim.episode(T, actions=acts, traj_obs=t_obs, traj_reward=t_rew, write_state=False)
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(20): sim.episode(T, actions=acts, traj_obs=t_obs, traj_reward=t_rew, write_state=False)
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 20)
    out[dt] = round(best, 5)
print(json.dumps(out))
