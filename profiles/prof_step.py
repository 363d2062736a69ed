"""Runs the single-period step kernel (VecEnv mode) at an HBM-honest size for ncu: 4 Mi envs, int32,
centralized complex scenario, with the row-major f32 observation output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200 import uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, 100, True)
sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
g = torch.Generator(device="cuda"); g.manual_seed(1)
sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device="cuda", dtype=torch.int32))
sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device="cuda", dtype=torch.int32))
acts = torch.randint(0, 17, (n, 4), generator=g, device="cuda", dtype=torch.int32)
sim.reset()
for t in range(8):
    sim.step(acts)
torch.cuda.synchronize()
print("ok")
