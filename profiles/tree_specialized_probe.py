"""General tree kernels vs a library specialised for the network (multi_echelon_drl_b200/specialize.py): fused rollout,
65,536 envs x 100 periods, trajectory out; bit-equality of every output, then timing.  The specialised libraries must
have been built beforehand (python profiles/tree_specialized_probe.py build  - in the build container)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_echelon_drl_b200 import specialize
from multi_echelon_drl_b200.spec import tree_spec

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
SPECS = {"tree, 2 levels, 2 sources": tree, "star, 3 retailers": star, "8 nodes: factory, annex, hub, 5 retailers": big}
if os.environ.get("PROBE_ONLY"):
    SPECS = {k: v for k, v in SPECS.items() if os.environ["PROBE_ONLY"] in k}

if len(sys.argv) > 1 and sys.argv[1] == "build":
    for name, sp in SPECS.items():
        print(name, specialize.build(sp, "int32", force=bool(os.environ.get("SN_SPEC_EXTRA_FLAGS"))))
    sys.exit(0)

import numpy as np, torch
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
n, T = 65536, 100
for name, sp in SPECS.items():
    res = {}
    for mode in ("general", "specialised"):
        os.environ["SN_NO_SPECIALIZED"] = "1" if mode == "general" else "0"
        specialize._loaded.clear()
        sim = BatchedSupplyNetwork(sp, n)
        assert (sim.fast is not None) == (mode == "specialised"), mode
        init, dem = bulk_episode_inputs(sp, n, np.random.RandomState(0))
        sim.load_episode(init, dem)
        acts = torch.as_tensor(np.random.RandomState(1).randint(-1, 18, (T, n, sp.n_agents)).astype(np.int32)).cuda()
        obs = torch.empty((T, n, sp.obs_dim), dtype=torch.float32, device="cuda")
        rew = torch.empty((T, n), dtype=torch.float32, device="cuda")
        for _ in range(3):
            sim.episode(T, actions=acts, traj_obs=obs, traj_reward=rew, write_state=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(20):
            sim.episode(T, actions=acts, traj_obs=obs, traj_reward=rew, write_state=False)
        b.record(); torch.cuda.synchronize()
        # step kernel of the same library, a few periods
        sim.reset()
        steps = [tuple(x.clone() for x in sim.step(acts[t])[:2]) for t in range(5)]
        res[mode] = dict(ms=a.elapsed_time(b) / 20, obs=obs.clone(), rew=rew.clone(), ret=sim.ep_return.clone(), state=sim.state.clone(), steps=steps)
    g, s = res["general"], res["specialised"]
    same = all(torch.equal(g[k], s[k]) for k in ("obs", "rew", "state", "ret")) and all(
        torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) for x, y in zip(g["steps"], s["steps"]))
    print(f"{name:28s} general {g['ms']:.4f} ms   specialised {s['ms']:.4f} ms   bit-identical: {same}")
