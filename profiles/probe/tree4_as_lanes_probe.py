"""VERDICT r1 item 6 asked to try the lane-per-node decomposition with FOUR lanes per env for trees of up to four
nodes against the register-resident tree kernel.  Timing only (random state words; the kernels' time does not depend on
the data): the bench's two-level tree, 65,536 envs x 100 periods, streamed actions, trajectory out.
  python profiles/probe/tree4_as_lanes_probe.py                         -> register-resident general kernel
  SUPPLYNET_LIB=.../variant_l4.so SN_TREE_AS_NET=1 python ...          -> four lanes per env (experimental build)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["SN_NO_SPECIALIZED"] = "1"
import torch
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.spec import tree_spec

tree = tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20),
                  dict(name="dist_b", demand=(0, 5), target=14), dict(name="maker", target=25)],
                 [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)],
                 ["shop", "dist_a", "dist_b", "maker"], True)
n, T = 65536, 100
sim = BatchedSupplyNetwork(tree, n, "cuda", "int32")
sim.init.random_(0, 9)
sim.demand.random_(0, 9)
acts = torch.randint(0, 17, (T, n, 4), device="cuda", dtype=torch.int32)
traj = torch.empty((T, n, sim.obs_dim), dtype=torch.float32, device="cuda")
for _ in range(3):
    sim.episode(T, actions=acts, traj_obs=traj)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    sim.episode(T, actions=acts, traj_obs=traj)
e.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(os.environ.get("SUPPLYNET_LIB", "libsupplynet.so")), "as_net": os.environ.get("SN_TREE_AS_NET", "0"),
                  "state_words": sim.state_words, "ms": round(s.elapsed_time(e) / 10, 4)}))
