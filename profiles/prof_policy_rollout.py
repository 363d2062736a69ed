"""Runs only the in-kernel base-stock rollout (config-3 mode) a few times, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs, T
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork, BasePolicyParams
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec, init, demand, actions = make_inputs(n, 1)
sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
sim.load_episode(init, demand)
pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
for _ in range(6):
    sim.episode(T, policy=pol, write_state=False)
torch.cuda.synchronize()
print("ok")
