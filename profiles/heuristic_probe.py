"""Bandwidth of the standalone base-stock kernel (sn_heuristic): obs [n,28] f32 in, actions [n,4] f32 out."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200.simulator import BasePolicyParams, heuristic_actions
pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
res = {}
for n in (65536, 4 * 1024 * 1024):
    obs = torch.randint(0, 30, (n, 28), device="cuda").float()
    out = torch.empty((n, 4), device="cuda")
    for _ in range(5): heuristic_actions(pol, obs, out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): heuristic_actions(pol, obs, out)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    res[n] = {"us": ms * 1e3, "GBs": n * 128 / ms / 1e6, "envs_per_s": n / ms * 1e3}
print(json.dumps(res))
