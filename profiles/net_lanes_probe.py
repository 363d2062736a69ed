"""General lane-per-node kernels on trees of five to eight nodes: 65,536 envs x 100 periods, streamed actions,
trajectory out (the bench's rollout shape), per node count.  usage: python profiles/net_lanes_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SN_NO_SPECIALIZED"] = "1"
import numpy as np
import torch
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.spec import tree_spec
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs

retail = [dict(name=f"r{i}", demand=(0, 8 - i), target=19 - 2 * i) for i in range(1, 6)]
out = {}
for nf in (5, 6, 7, 8):
    k = nf - 3 if nf < 8 else 5                      # retailers under the hub
    nodes = [dict(name="hub", target=60)] + retail[:k] + [dict(name="factory", target=40)] + ([dict(name="annex", demand=(0, 5), target=12)] if nf - k - 2 else [])
    arcs = [("ext", "factory", 1, 2)] + ([("factory", "annex", 1, 3)] if nf - k - 2 else []) + [("factory", "hub", 2, 2)] + [("hub", f"r{i}", 1 + i % 2, 1 + i % 2) for i in range(1, k + 1)]
    spec = tree_spec(nodes, arcs, [n["name"] for n in nodes], True)
    assert spec.n_facilities == nf, (nf, spec.n_facilities)
    n, T = 65536, 100
    sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
    init, dem = bulk_episode_inputs(spec, n, np.random.RandomState(nf))
    sim.load_episode(init, dem)
    acts = torch.randint(0, 17, (T, n, nf), device="cuda", dtype=torch.int32)
    traj = torch.empty((T, n, sim.obs_dim), dtype=torch.float32, device="cuda")
    for _ in range(3):
        sim.episode(T, actions=acts, traj_obs=traj)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        sim.episode(T, actions=acts, traj_obs=traj)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    out[f"nodes_{nf}"] = {"ms": round(ms, 4), "env_steps_per_s": round(n * T / ms * 1e3), "lanes_per_env": 5 if nf == 5 else 6 if nf == 6 else 8}
print(json.dumps(out))
