"""Time of the fused rollout (65,536 envs x 100 periods, trajectory out) on the three kernel families:
tuned beer-game chain, generic lead times, distribution tree.  CUDA events, 20 launches after 3 warm-ups."""
import sys
from dataclasses import replace

import numpy as np
import torch

sys.path.insert(0, ".")
from multi_echelon_drl_b200 import uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.spec import tree_spec
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs

n, T = 65536, 100
tree = tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20), dict(name="dist_b", demand=(0, 5), target=14),
                  dict(name="maker", target=25)],
                 [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)],
                 ["shop", "dist_a", "dist_b", "maker"], True)
star = tree_spec([dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=12), dict(name="r3", demand=(0, 3), target=9),
                  dict(name="hub", target=40)],
                 [("ext", "hub", 2, 3), ("hub", "r3", 1, 1), ("hub", "r1", 2, 2), ("hub", "r2", 2, 1)], ["r1", "r2", "r3", "hub"], True)
big = tree_spec([dict(name="hub", target=60), dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=14),
                 dict(name="r3", demand=(0, 4), target=10), dict(name="r4", demand=(0, 3), target=8), dict(name="r5", demand=(0, 2), target=6),
                 dict(name="factory", target=40), dict(name="annex", demand=(0, 5), target=12)],
                [("ext", "factory", 1, 2), ("factory", "annex", 1, 3), ("factory", "hub", 2, 2), ("hub", "r3", 2, 1), ("hub", "r1", 2, 2),
                 ("hub", "r5", 3, 1), ("hub", "r2", 1, 1), ("hub", "r4", 1, 2)],
                ["hub", "r1", "r2", "r3", "r4", "r5", "factory", "annex"], True)
five = tree_spec([dict(name="kiosk", demand=(0, 3), target=8), dict(name="shop_a", demand=(0, 8), target=19), dict(name="shop_b", demand=(0, 5), target=16),
                  dict(name="dist", target=35), dict(name="maker", target=30)],
                 [("ext", "maker", 1, 2), ("maker", "dist", 2, 2), ("dist", "shop_b", 2, 2), ("dist", "shop_a", 2, 2), ("shop_b", "kiosk", 1, 1)],
                 ["kiosk", "shop_a", "shop_b", "dist", "maker"], True)
specs = {"beer game (tuned)": uniform_spec(None, True), "other lead times (generic)": replace(uniform_spec(None, True), info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1)),
         "tree, 2 levels, 2 sources": tree, "star, 3 retailers (3 fill rounds)": star,
         "5-node tree (lane-per-node kernels)": five, "8-node tree (lane-per-node kernels)": big}
for name, sp in specs.items():
    sim = BatchedSupplyNetwork(sp, n)
    init, dem = bulk_episode_inputs(sp, n, np.random.RandomState(0))
    sim.load_episode(init, dem)
    acts = torch.randint(0, 17, (T, n, sp.n_agents), device="cuda", dtype=torch.int32)
    obs = torch.empty((T, n, sp.obs_dim), dtype=torch.float32, device="cuda")
    rew = torch.empty((T, n), dtype=torch.float32, device="cuda")
    for _ in range(3):
        sim.episode(T, actions=acts, traj_obs=obs, traj_reward=rew, write_state=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        sim.episode(T, actions=acts, traj_obs=obs, traj_reward=rew, write_state=False)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    nsrc = sp.n_sources
    bytes_ = n * T * (4 * sp.n_agents + 4 * nsrc + 4 * sp.obs_dim + 4)
    print(f"{name:36s} {ms:.4f} ms  {n * T / ms / 1e6:7.2f} G env-steps/s  {bytes_ / ms / 1e6:7.1f} GB/s")
