"""Distribution trees on the CUDA path (SN_TOPO_TREE, supplynet_tree.cuh) vs trajectories recorded from the
reference's own classes (tests/golden/trees.npz), vs the tree oracle on batches, and vs the chain kernels on
networks that are chains.  Integer quantities: bit-exact; float quantities: 1e-6 relative."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_oracle_trees import BIG_TREE_CASES, TREE_CASES, tree_oracle_rollout, tree_product_spec  # noqa: E402


def _sim(case, n, dtype="int32", lib="general"):
    """lib = "general": the run-time-topology kernels of libsupplynet.so; "specialised": the library compiled for this
    very network (multi_echelon_drl_b200/specialize.py; __graft_entry__.build() prebuilds the int32 ones of the
    golden topologies), which BatchedSupplyNetwork picks up on its own when it exists."""
    import os
    from multi_echelon_drl_b200 import specialize
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    os.environ["SN_NO_SPECIALIZED"] = "0" if lib == "specialised" else "1"
    try:
        sim = BatchedSupplyNetwork(tree_product_spec(case), n, "cuda", dtype)
    finally:
        os.environ.pop("SN_NO_SPECIALIZED", None)
    if lib == "specialised":
        assert sim.fast is not None, f"no specialised library for {case['id']}: run __graft_entry__.build()"
        assert os.path.basename(specialize.lib_path(sim.spec, dtype)).startswith("libsupplynet_tree_")
    else:
        assert sim.fast is None
    return sim


@pytest.mark.parametrize("lib", ["general", "specialised"])
@pytest.mark.parametrize("case", TREE_CASES + BIG_TREE_CASES, ids=[c["id"] for c in TREE_CASES + BIG_TREE_CASES])
def test_tree_step_and_rollout_match_reference_golden(case, lib):
    """Up to four nodes: register-resident tree kernels (64 state words); five to eight: lane-per-node kernels (128
    state words).  Each also through the library specialised for the network (register-resident for every size: eight
    facilities per thread for the larger ones, same 128-word state buffer)."""
    sim = _sim(case, 1, lib=lib)
    N = len(case["names"])
    half = 8 if N > 4 else 4                      # rows of the per-facility cost output: holding[half], backorder[half]
    assert sim.state_words == (128 if N > 4 else 64) and sim.obs_dim == case["obs0"].shape[0]
    sim.load_episode(case["init"][None], case["demand_table"][None])
    assert np.array_equal(sim.reset().cpu().numpy()[0], case["obs0"])
    r64 = torch.zeros(1, dtype=torch.float64, device="cuda")
    cost = torch.zeros((2 * half, sim.ld), dtype=torch.float64, device="cuda")
    for t in range(100):
        obs, _, done = sim.step(torch.as_tensor(case["actions"][t][None]), reward64_out=r64, cost_out=cost)
        assert np.array_equal(obs.cpu().numpy()[0], case["obs"][t]), t
        assert r64.item() == case["rew"][t], t
        c = cost[:, 0].cpu().numpy()
        assert np.array_equal(c[:N], case["ref_holding"][t]) and np.array_equal(c[half:half + N], case["ref_backorder"][t]), t
    assert done
    out = sim.episode(100, actions=torch.as_tensor(case["actions"][:, None, :]), record_obs=True, record_reward=True)
    assert np.array_equal(out["traj_obs"].cpu().numpy()[:, 0], case["obs"])
    assert sim.ep_return.item() == case["rew"].sum()


@pytest.mark.parametrize("cid,dtype", [(3, "int32"), (20, "int32"), (37, "int32"), (55, "int32"), (70, "int32"), (88, "int32"),
                                       (37, "float32"), (55, "float64"),
                                       (1000, "int32"), (1013, "int32"), (1024, "int32"), (1037, "int32"), (1049, "int32"),
                                       (1001, "float32"), (1024, "float64")])
@pytest.mark.parametrize("lib", ["general", "specialised"])
def test_tree_batch_vs_oracle(cid, dtype, lib):
    """3,001 envs (ragged) x 100 with bulk-drawn initial states and per-source demand streams: step kernel, fused
    rollout, rollout resumed mid-episode, in-kernel base-stock policy."""
    from multi_echelon_drl_b200.simulator import BasePolicyParams, heuristic_actions
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    case = BIG_TREE_CASES[cid - 1000] if cid >= 1000 else TREE_CASES[cid]      # 1000+: five to eight nodes
    if lib == "specialised" and dtype != "int32":
        pytest.skip("prebuilt specialised libraries are int32")
    n = 3001
    sim = _sim(case, n, dtype, lib=lib)
    init, demand = bulk_episode_inputs(sim.spec, n, np.random.RandomState(cid))
    rs = np.random.RandomState(cid + 1)
    acts = rs.randint(-1, 19, (100, n, len(case["agents"])))
    if dtype != "int32":
        acts = acts + rs.rand(*acts.shape).round(3)
    import oracle.network_np as NP
    if dtype == "int32":
        obs0, obs, rew = tree_oracle_rollout(case, init, demand, acts)
    else:
        from test_oracle_trees import oracle_spec
        orc = NP.TreeOracle(oracle_spec(case), n, dtype=np.float64)
        N = len(case["names"])
        from test_oracle_trees import init_layout
        _, so0, sh0 = init_layout(N)
        obs0 = orc.reset(init[:, :N].astype(np.float64), [init[:, so0 + 4 * k: so0 + 4 * k + case["li"][k]].astype(np.float64) for k in range(N)],
                         [init[:, sh0 + 4 * k: sh0 + 4 * k + case["ls"][k]].astype(np.float64) for k in range(N)], demand)
        obs, rew = [], []
        for t in range(100):
            o, r, _ = orc.step(acts[t]); obs.append(o); rew.append(r)
        obs, rew = np.stack(obs), np.stack(rew)
    sim.load_episode(init, demand)
    # float quantities: BASELINE.json's 1e-6 relative - to the scale of the trajectory (max |obs| / max |reward|), as for
    # the chains (test_gpu_parity.py); float64 kernels agree with the float64 restatement to the float32 output's rounding
    oscale, rscale = max(1.0, float(np.abs(obs).max())), max(1.0, float(np.abs(rew).max()))
    tol = 1e-7 if dtype == "float64" else 1e-6

    def close(a, b, scale=oscale):
        if dtype == "int32":
            return np.array_equal(a, b)
        return np.abs(np.asarray(a, np.float64) - b).max() <= tol * scale
    assert close(sim.reset().cpu().numpy(), obs0)
    r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(100):
        o, _, _ = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
        assert close(o.cpu().numpy(), obs[t]), t
        assert close(r64.cpu().numpy(), rew[t], rscale), t
    stepped_return = sim.ep_return.clone()
    out = sim.episode(100, actions=torch.as_tensor(acts), record_obs=True)
    assert close(out["traj_obs"].cpu().numpy(), obs)
    assert close(sim.ep_return.cpu().numpy(), rew.sum(0), max(1.0, float(np.abs(rew.sum(0)).max())))
    if dtype == "int32":
        assert torch.equal(sim.ep_return, stepped_return)
        # resume: 37 periods stepped, the rest as one rollout launch
        sim.reset()
        for t in range(37):
            sim.step(torch.as_tensor(acts[t]))
        out = sim.rollout(63, actions=torch.as_tensor(acts[37:]), record_obs=True)
        assert np.array_equal(out["traj_obs"].cpu().numpy(), obs[37:]) and torch.equal(sim.ep_return, stepped_return)
        # in-kernel base-stock policy == heuristic kernel + step kernel
        if case["glob"] and len(case["agents"]) == len(case["names"]) and list(case["agents"]) == sorted(case["agents"]):
            pol = BasePolicyParams(case["target"], 0, 16, "a")
            sim.episode(100, policy=pol, write_state=False)
            fused = sim.ep_return.clone()
            o = sim.reset()
            for t in range(100):
                o, _, _ = sim.step(heuristic_actions(pol, o))
            assert torch.equal(fused, sim.ep_return)


@pytest.mark.parametrize("li,ls,agents,glob,rule", [((2, 2, 2, 1), (2, 2, 2, 2), (0, 1, 2, 3), True, "a"),
                                                    ((3, 1, 2, 2), (2, 4, 3, 1), (2, 1), False, "d+a"),
                                                    ((1, 2, 2), (2, 2, 3), (1,), True, "a")])
def test_chain_through_tree_kernels_equals_chain_kernels(li, ls, agents, glob, rule):
    """A serial chain described as a tree runs on the tree kernels and gives what the chain kernels give."""
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    nf, n = len(li), 2049
    names = ["retailer", "wholesaler", "distributor", "manufacturer"][:nf]
    u = uniform_spec(None, glob, 100)
    chain = replace(u, n_facilities=nf, agents=agents, info_leadtime=li, ship_leadtime=ls, holding_cost=u.holding_cost[:nf],
                    backorder_cost=u.backorder_cost[:nf], coplayer_target=u.coplayer_target[:nf])
    nodes = [dict(name=names[k], holding_cost=0.5, backorder_cost=1.0, target=u.coplayer_target[k], demand=(0, 8) if k == 0 else None)
             for k in range(nf)]
    arcs = [(names[k + 1] if k + 1 < nf else "external_supplier", names[k], li[k], ls[k]) for k in range(nf - 1, -1, -1)]
    tree = tree_spec(nodes, arcs, [names[a] for a in agents], glob)
    if rule == "d+a":
        chain, tree = chain.with_rule("d+a", -8, 0, 16), tree.with_rule("d+a", -8, 0, 16)
    a, b = BatchedSupplyNetwork(chain, n), BatchedSupplyNetwork(tree, n)
    assert b.state_words == 64
    init, dem = bulk_episode_inputs(tree, n, np.random.RandomState(4))
    init_chain = init
    if a.init_words == 19:
        cols = list(range(4)) + [4 + 4 * k + j for k in range(4) for j in range(li[k])] + [20 + 4 * k + j for k in range(4) for j in range(ls[k])]
        init_chain = init[:, cols]
    a.load_episode(init_chain, dem[:, 0]); b.load_episode(init, dem)
    acts = torch.as_tensor(np.random.RandomState(5).randint(-1, 19, (100, n, len(agents))))
    oa = a.episode(100, actions=acts, record_obs=True, record_reward=True)
    ob = b.episode(100, actions=acts, record_obs=True, record_reward=True)
    assert torch.equal(oa["traj_obs"], ob["traj_obs"]) and torch.equal(oa["traj_reward"], ob["traj_reward"])
    assert torch.equal(a.ep_return, b.ep_return)


def test_vecenv_on_a_tree_and_statistics():
    """SupplyNetVecEnv on a two-level tree with two demand sources: device-drawn episodes, auto-reset, on-order
    invariant, policy episodes; per-source demand ranges are honoured."""
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    case = TREE_CASES[[c["topo"] for c in TREE_CASES].index("two_level")]
    assert case["glob"] is False
    spec = tree_product_spec(dict(case, agents=[0, 1, 2, 3], glob=True, rule="a"))
    env = SupplyNetVecEnv(spec, 1000, seed=2, dtype="int32")
    obs = env.reset()
    assert obs.shape == (1000, 28) and env.action_space.shape == (4,)
    dem = env.sim.demand[:, :, :1000]
    assert dem.shape[:2] == (103, 2) and int(dem[:, 0].max()) == 8 and int(dem[:, 1].max()) == 5 and int(dem.min()) == 0
    s = env.sim.state[:, :1000].to(torch.int64)
    for k in range(4):                                   # on_order == sum(pipeline) on every arc
        on_order = s[7 * k + 3:7 * k + 7].sum(0)
        pipe = s[28 + 3 * k:31 + 3 * k].sum(0) + s[40 + 4 * k:44 + 4 * k].sum(0) + s[56 + k]
        assert torch.equal(on_order, pipe)
    for _ in range(230):
        obs, rew, dones, infos = env.step(torch.full((1000, 4), 5, device="cuda"))
    assert torch.isfinite(obs).all() and env.episodes_done == 2000
    stats = env.run_policy_episodes(BasePolicyParams(case["target"], 0, 16, "a"), n_episodes=3).cpu().numpy()
    assert stats[0] == 3000 * 100 and stats[10] == 3000 and stats[11] < 0 and stats[11] == stats[1]


def test_single_env_from_yaml_replays_the_reference_episode():
    """make_env_from_config on the packaged tree YAML: `np.random.seed(s); env.reset(); env.step(a)...` gives the
    observations and rewards the reference gave for the same seed and actions (golden case of the same network)."""
    import os
    import multi_echelon_drl_b200 as pkg
    from multi_echelon_drl_b200.envs import make_env_from_config
    path = os.path.join(os.path.dirname(pkg.__file__), "network_configs", "distribution_tree.yaml")
    for glob, rule, agents in ((True, "a", [1, 2]), (False, "a", [3])):
        c = [c for c in TREE_CASES if c["topo"] == "two_level" and c["agents"] == agents and c["glob"] == glob and c["rule"] == rule][0]
        env = make_env_from_config(path, [c["names"][a] for a in agents], glob)
        assert env.sn.order_sequence == c["order_sequence"] + ["external_supplier"]
        np.random.seed(c["seed"])
        obs = env.reset()
        assert np.array_equal(obs, c["obs0"])
        for t in range(100):
            obs, r, done, info = env.step(c["actions"][t])
            assert np.array_equal(obs, c["obs"][t]) and r == c["rew"][t], t
        assert done
        st = env.sn.get_state("maker")
        assert st["on_hand"] == (c["obs"][-1][21] if glob else c["obs"][-1][0])
        with pytest.raises(ValueError):
            env.step(c["actions"][0])
    venv = make_env_from_config(path, ["shop", "dist_a", "dist_b", "maker"], True, n_envs=256, dtype="int32")
    assert venv.reset().shape == (256, 28)


def test_random_trees_vs_oracle():
    """40 random trees, 16 of them with five to eight nodes (structure, names, arc order, lead times, sources, agent sets, rules): fused episode and
    stepwise kernels vs the tree oracle on 257 envs each, bit-exact."""
    from test_oracle_trees import random_tree_case, random_tree_inputs
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    rs = np.random.RandomState(31)
    for i in range(40):
        c = random_tree_case(rs, None if i < 24 else 5 + i % 4)           # the last 16: five to eight nodes
        n = 257
        init, dem, acts = random_tree_inputs(c, n, rs)
        obs0, obs, rew = tree_oracle_rollout(c, init, dem, acts)
        sim = BatchedSupplyNetwork(tree_product_spec(c), n, "cuda", "int32")
        sim.load_episode(init, dem)
        o0 = torch.empty((n, sim.obs_dim), dtype=torch.float32, device="cuda")
        out = sim.episode(100, actions=torch.as_tensor(acts), record_obs=True, record_reward=True, obs0_out=o0)
        assert np.array_equal(o0.cpu().numpy(), obs0), (i, c)
        assert np.array_equal(out["traj_obs"].cpu().numpy(), obs), (i, c)
        assert np.array_equal(sim.ep_return.cpu().numpy(), rew.sum(0)), (i, c)
        if i % 4 == 0:
            assert np.array_equal(sim.reset().cpu().numpy(), obs0)
            r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
            for t in range(100):
                o, _, _ = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
                assert np.array_equal(o.cpu().numpy(), obs[t]) and np.array_equal(r64.cpu().numpy(), rew[t]), (i, t, c)


def test_vecenv_and_heuristic_on_an_eight_node_tree():
    """hub with five retailers under a factory that also feeds an annex (8 nodes, 6 demand sources): VecEnv with
    device-drawn episodes, the wide heuristic kernel vs numpy, in-kernel policy episodes vs heuristic + step."""
    from multi_echelon_drl_b200.simulator import BasePolicyParams, heuristic_actions
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    import oracle.beergame_np as O
    case = BIG_TREE_CASES[[c["topo"] for c in BIG_TREE_CASES].index("eight_hub")]
    spec = tree_product_spec(dict(case, agents=list(range(8)), glob=True, rule="a"))
    env = SupplyNetVecEnv(spec, 777, seed=4, dtype="int32")
    obs = env.reset()
    assert obs.shape == (777, 56) and env.action_space.shape == (8,) and env.sim.state_words == 128 and env.sim.init_words == 72
    pol = BasePolicyParams(case["target"], 0, 16, "a")
    a = heuristic_actions(pol, obs)
    assert np.array_equal(a.cpu().numpy(), O.base_stock_action(obs.cpu().numpy(), case["target"], 0, 16, "a", n_facilities=8).astype(np.float32))
    s = env.sim.state[:, :777].to(torch.int64)
    for k in range(8):                                   # on_order == sum(pipeline) on every arc
        b = 16 * k
        assert torch.equal(s[b + 3:b + 7].sum(0), s[b + 7:b + 10].sum(0) + s[b + 10:b + 14].sum(0) + s[b + 14])
    total = torch.zeros(777, dtype=torch.float64, device="cuda")
    for _ in range(100):
        obs, rew, dones, infos = env.step(heuristic_actions(pol, obs))
        total += rew.double()
    assert dones.all() and torch.equal(infos.episode_return, total)
    stepped = total.clone()
    env.sim.episode(100, policy=pol, write_state=False)            # same episode tables: the auto-reset drew new ones
    stats = env.run_policy_episodes(pol, n_episodes=2).cpu().numpy()
    assert stats[0] == 2 * 777 * 100 and stats[10] == 2 * 777 and stats[11] < 0 and stats[11] == stats[1]
    assert not stats[2:10].any()                                    # per-facility sums are not kept for these networks
    assert (stepped < 0).all()


def test_big_tree_rewards_with_cost_coefficients_that_round():
    """Cost coefficients off the 2^-10 grid (0.3, 0.7, 1.1): the lane-per-node kernels must then add the per-node terms
    strictly in node order, like SupplyNetwork.get_cost; rewards stay bit-identical to the oracle's float64 sums."""
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    base = BIG_TREE_CASES[[c["topo"] for c in BIG_TREE_CASES].index("seven_two_levels")]
    case = dict(base, holding=[0.3, 0.7, 1.1, 0.3, 0.123, 0.9, 0.45], backorder=[1.1, 0.7, 2.3, 0.3, 1.7, 0.6, 3.3])
    n = 1025
    sim = BatchedSupplyNetwork(tree_product_spec(case), n, "cuda", "int32")
    init, demand = bulk_episode_inputs(sim.spec, n, np.random.RandomState(8))
    acts = np.random.RandomState(9).randint(-1, 19, (100, n, len(case["agents"])))
    obs0, obs, rew = tree_oracle_rollout(case, init, demand, acts)
    sim.load_episode(init, demand)
    sim.reset()
    r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(100):
        o, _, _ = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
        assert np.array_equal(o.cpu().numpy(), obs[t]) and np.array_equal(r64.cpu().numpy(), rew[t]), t
    sim.episode(100, actions=torch.as_tensor(acts))
    total = np.zeros(n)
    for t in range(100):
        total = total + rew[t]                       # the kernel accumulates the episode return period by period
    assert np.array_equal(sim.ep_return.cpu().numpy(), total)


def test_single_env_on_big_trees_replays_the_reference_episodes():
    """The gym-style single env on five- to eight-node networks: `np.random.seed(s); env.reset(); env.step(a)...`
    reproduces the reference's observations and rewards (golden cases), and `sn.get_state(name)` reads any node."""
    import warnings
    from multi_echelon_drl_b200.gym_env import BeerGameEnv
    for cid in (0, 7, 13, 26, 41, 55):
        c = BIG_TREE_CASES[cid]
        env = BeerGameEnv(tree_product_spec(c), "cuda", "int32")
        assert env.sn.order_sequence == c["order_sequence"] + ["external_supplier"]
        np.random.seed(c["seed"])
        assert np.array_equal(env.reset(), c["obs0"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)          # golden actions include negative orders
            for t in range(100):
                obs, r, done, _ = env.step(c["actions"][t])
                assert np.array_equal(obs, c["obs"][t]) and r == c["rew"][t], (cid, t)
        assert done
        last = c["names"][-1]
        st = env.sn.get_state(last)
        if c["glob"]:
            assert st["on_hand"] == c["obs"][-1][7 * (len(c["names"]) - 1)]
        assert set(st) == {"on_hand", "unfilled_demand", "latest_demand"} | {f"unreceived_pipeline_{i}" for i in range(4)}


@pytest.mark.parametrize("nf", [5, 6, 7])
def test_lane_packed_trees_at_warp_and_block_boundaries(nf):
    """Trees of five / six nodes take five / six lanes per env: six / five envs share a warp, 24 / 20 a block, two lanes
    of every warp idle (seven and eight nodes: eight lanes, four envs).  Batch sizes around those boundaries, fused
    episode + stepwise kernels + observe vs the tree oracle, bit-exact; the buffers' tails stay untouched."""
    from test_oracle_trees import random_tree_case, random_tree_inputs
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    rs = np.random.RandomState(500 + nf)
    c = random_tree_case(rs, nf)
    for n in (1, 4, 5, 6, 7, 19, 20, 21, 23, 24, 25, 47, 49):
        init, dem, acts = random_tree_inputs(c, n, rs)
        obs0, obs, rew = tree_oracle_rollout(c, init, dem, acts)
        sim = BatchedSupplyNetwork(tree_product_spec(c), n, "cuda", "int32")
        assert sim.fast is None                      # a random topology: the general lane-per-node kernels
        sim.load_episode(init, dem)
        T = obs.shape[0]
        traj = torch.full((T * n * sim.obs_dim + 64,), -7.0, dtype=torch.float32, device="cuda")     # canary tail
        out = sim.episode(T, actions=torch.as_tensor(acts), traj_obs=traj[: T * n * sim.obs_dim].view(T, n, sim.obs_dim),
                          record_reward=True)
        assert np.array_equal(out["traj_obs"].cpu().numpy(), obs), (nf, n)
        assert (traj[T * n * sim.obs_dim:] == -7.0).all(), (nf, n)
        assert np.array_equal(out["traj_reward"].cpu().numpy().astype(np.float64), rew.astype(np.float32).astype(np.float64)), (nf, n)
        assert np.array_equal(sim.ep_return.cpu().numpy(), rew.sum(0)), (nf, n)
        assert np.array_equal(sim.reset().cpu().numpy(), obs0)
        for t in range(5):
            o, _, _ = sim.step(torch.as_tensor(acts[t]))
            assert np.array_equal(o.cpu().numpy(), obs[t]), (nf, n, t)
            assert np.array_equal(sim.observe(torch.empty_like(o)).cpu().numpy(), obs[t])
