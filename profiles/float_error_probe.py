"""Measured float errors (what the tolerances in tests/test_gpu_parity.py are set from): float32 / float64 kernels vs
(a) the float golden trajectories recorded from the reference (which itself computes those in float32 scalars under
numpy >= 2) and (b) the float64 oracle on a 4,096-env batch.  Errors relative to the state scale max|obs| and, for
rewards, to max|reward|."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden_rollouts
from helpers import oracle_rollout
from test_gpu_parity import _make, _seeded_batch
out = {}
for dt in ("float32", "float64"):
    worst_o = worst_r = 0.0
    for case in [c for c in load_golden_rollouts() if c["kind"] != "int"]:
        sim = _make(case["scenario"], case["agents"], case["global_observable"], case["rule"], 100, 1, dt)
        sim.load_episode(case["init"][None], case["demand"][None]); sim.reset()
        r64 = torch.zeros((1,), dtype=torch.float64, device="cuda")
        scale = max(1.0, np.abs(case["obs"]).max()); rs = max(1.0, np.abs(case["rew"]).max())
        for t in range(case["actions"].shape[0]):
            obs, rew, done = sim.step(torch.as_tensor(case["actions"][t][None]), reward64_out=r64)
            worst_o = max(worst_o, float(np.abs(obs.cpu().numpy()[0].astype(np.float64) - case["obs"][t]).max() / scale))
            worst_r = max(worst_r, abs(r64.item() - case["rew"][t]) / rs)
    out[f"golden_{dt}"] = {"obs_err_over_scale": worst_o, "reward_err_over_scale": worst_r}
n, T = 4096, 100
init, demand = _seeded_batch("uniform", n, T, 44)
acts = (np.random.RandomState(4).rand(T, n, 4) * 16).astype(np.float32)
ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts.astype(np.float64), dtype=np.float64)
scale = float(np.abs(ref["obs"]).max()); rs = float(np.abs(ref["rew"]).max())
for dt in ("float32", "float64"):
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, dt)
    sim.load_episode(init, demand); sim.reset()
    o = sim.rollout(T, actions=torch.as_tensor(acts), record_obs=True, record_reward=True)
    eo = np.abs(o["traj_obs"].cpu().numpy().astype(np.float64) - ref["obs"]).max() / scale
    er = np.abs(o["traj_reward"].cpu().numpy().astype(np.float64) - ref["rew"]).max() / rs
    eret = np.abs(sim.ep_return.cpu().numpy() - ref["rew"].sum(0)).max() / np.abs(ref["rew"].sum(0)).max()
    out[f"oracle_{dt}"] = {"obs_err_over_scale": float(eo), "reward_err_over_scale": float(er), "ep_return_rel": float(eret), "scale": scale}
print(json.dumps(out, indent=1))
