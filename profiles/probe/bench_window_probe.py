"""Where the fixed cost of a short timed window goes: completion time of every launch of a K-launch window of the
config-2 episode kernel (bench.py's device_step), relative to the event recorded right after the synchronize, and the
CPU time Python needs to issue one launch."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
n, T = 65536, 100
dev = torch.device("cuda", 0)
spec, init, demand, actions = bench.make_inputs(n, 1234)
sim = BatchedSupplyNetwork(spec, n, dev, "int32")
sim.track_stats = True
sim.load_episode(torch.from_numpy(init).pin_memory(), torch.from_numpy(demand).pin_memory())
d_actions = torch.from_numpy(actions).to(dev)
traj_obs = torch.empty((T, n, 28), dtype=torch.float32, device=dev)
traj_rew = torch.empty((T, n), dtype=torch.float32, device=dev)
step = lambda: sim.episode(T, actions=d_actions, traj_obs=traj_obs, traj_reward=traj_rew, write_state=False)
out = {}
for _ in range(5): step()
torch.cuda.synchronize()
for K in (3, 20, 50):
    res = []
    for rep in range(3):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(K): step()
        e.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        res.append((round(s.elapsed_time(e) / K, 5), round((t1 - t0) * 1e6 / K, 1)))
    out[f"K{K}_ms_per_step__cpu_us_per_issue"] = res
# completion times inside one 20-launch window (one event after every launch)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); ev = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
s.record()
for i in range(20):
    step(); ev[i].record()
torch.cuda.synchronize()
ts = [s.elapsed_time(x) for x in ev]
out["completion_ms"] = [round(x, 4) for x in ts]
out["durations_ms"] = [round(ts[0], 4)] + [round(b - a, 4) for a, b in zip(ts, ts[1:])]
print(json.dumps(out))
