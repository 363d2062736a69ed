"""GPU tests of the VecEnv surface (SB3 duck type), registry ids, d+a wrapper and the heuristic
class, checked against the numpy oracle on per-env seeded streams (the reference's RNG order)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import oracle_rollout  # noqa: E402


def _oracle_episode(spec, scenario, seeds_streams, acts, rule="a"):
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs_from_streams
    init, demand = seeded_episode_inputs_from_streams(spec, seeds_streams)
    return oracle_rollout(scenario, spec.agents, spec.global_observable, rule, spec.max_episode_steps, init, demand, acts)


def test_vecenv_autoreset_matches_oracle_over_three_episodes():
    """numpy_seeded source: env i == reference env seeded with seed+i and reset repeatedly.
    Checks obs/reward every step, terminal_observation, Monitor-style episode info, auto-reset obs."""
    from multi_echelon_drl_b200 import register_envs as R
    n, T, seed = 37, 100, 500
    env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=seed,
                 episode_source="numpy_seeded")
    assert env.num_envs == n and env.observation_space.shape == (28,) and env.action_space.shape == (4,)
    assert float(env.action_space.high[0]) == 16.0
    streams = [np.random.RandomState(seed + i) for i in range(n)]
    rs = np.random.RandomState(0)
    obs = env.reset()
    for ep in range(3):
        acts = rs.randint(0, 17, (T, n, 4))
        ref = _oracle_episode(env.spec, "uniform", streams, acts)
        assert np.array_equal(obs.cpu().numpy(), ref["obs0"]), ep
        for t in range(T):
            obs, rew, dones, infos = env.step(torch.as_tensor(acts[t]).cuda())
            assert np.array_equal(rew.cpu().numpy(), ref["rew"][t].astype(np.float32))
            if t < T - 1:
                assert not dones.any() and infos[0] == {} and len(infos) == n
                assert np.array_equal(obs.cpu().numpy(), ref["obs"][t])
            else:
                assert dones.all()
                assert np.array_equal(infos.terminal_observation.cpu().numpy(), ref["obs"][t])
                i = 5
                assert infos[i]["episode"]["r"] == round(float(ref["rew"][:, i].sum()), 6)
                assert infos[i]["episode"]["l"] == T and infos[i]["TimeLimit.truncated"] is False
                assert np.array_equal(infos[i]["terminal_observation"].cpu().numpy(), ref["obs"][t][i])
                # obs is already the first observation of the next episode (DummyVecEnv auto-reset)
    assert env.episodes_done == 3 * n


def test_numpy_output_mode_and_decentralized_roles():
    """output='numpy' (pinned host staging, what SB3 consumes) on a decentralized Retailer env with
    local observation (config 1 semantics: BeerGameNormalRetailer-v0)."""
    from multi_echelon_drl_b200 import register_envs as R
    n, T, seed = 64, 100, 9
    env = R.make("BeerGameNormalRetailer-v0", n_envs=n, dtype="int32", seed=seed, episode_source="numpy_seeded",
                 output="numpy")
    assert env.observation_space.shape == (7,) and env.action_space.shape == (1,) and float(env.action_space.high[0]) == 20.0
    streams = [np.random.RandomState(seed + i) for i in range(n)]
    acts = np.random.RandomState(1).randint(0, 21, (T, n, 1))
    ref = _oracle_episode(env.spec, "normal", streams, acts)
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float32 and np.array_equal(obs, ref["obs0"])
    for t in range(T):
        obs, rew, dones, infos = env.step(acts[t])
        assert isinstance(rew, np.ndarray) and np.array_equal(rew, ref["rew"][t].astype(np.float32))
        if t < T - 1:
            assert np.array_equal(obs, ref["obs"][t]) and not dones.any()
    assert dones.all() and np.array_equal(infos.terminal_observation, ref["obs"][-1])
    assert np.allclose(infos.episode_return, ref["rew"].sum(0))


def test_numpy_outputs_survive_the_next_steps():
    """SB3 algorithms keep `self._last_obs = new_obs` across env.step() (DummyVecEnv hands out fresh arrays): what a
    step returned must not change under the caller when the env steps again.  The staging buffers rotate over
    kRing = 3 sets, so an array stays intact for the next two calls; wrong-shaped and fractional-for-int32 numpy
    actions are rejected like torch ones."""
    from multi_echelon_drl_b200 import register_envs as R
    n = 257
    env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=3, output="numpy")
    rs = np.random.RandomState(0)
    prev = env.reset()
    keep = prev.copy()
    history = []
    for t in range(6):
        obs, rew, dones, infos = env.step(rs.randint(0, 17, (n, 4)))
        assert np.array_equal(prev, keep), t            # the previous observation is untouched by this step
        assert not np.shares_memory(obs, prev)
        history.append((obs, obs.copy(), rew, rew.copy(), dones, dones.copy()))
        if len(history) >= 2:                           # ... and still is one step later
            o, oc, r, rc, d, dc = history[-2]
            assert np.array_equal(o, oc) and np.array_equal(r, rc) and np.array_equal(d, dc)
        prev, keep = obs, obs.copy()
    with pytest.raises(ValueError):
        env.step(np.zeros((n, 3)))
    with pytest.raises(TypeError):
        env.step(np.full((n, 4), 0.5))


def test_dplusa_registry_id_and_wrapper():
    """BeerGameUniformMultiFacilityDPlusA-v0 = wrap_action_d_plus_a(env, -8, 0, 16) (register_envs.py:8-15)."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.envs import make_beer_game_uniform_multi_facility
    from multi_echelon_drl_b200.utils.wrappers import wrap_action_d_plus_a
    n, T, seed = 33, 100, 77
    env = R.make("BeerGameUniformMultiFacilityDPlusA-v0", n_envs=n, dtype="int32", seed=seed, episode_source="numpy_seeded")
    assert env.spec.rule == "d+a" and env.spec.dpa_offset == -8 and env.observation_space.shape == (28,)
    env2 = wrap_action_d_plus_a(make_beer_game_uniform_multi_facility(
        agent_managed_facilities=["retailer", "wholesaler", "distributor", "manufacturer"], box_action_space=True,
        random_init=True, n_envs=n, dtype="int32", seed=seed, episode_source="numpy_seeded"), offset=-8, lb=0, ub=16)
    streams = [np.random.RandomState(seed + i) for i in range(n)]
    acts = np.random.RandomState(3).randint(0, 17, (T, n, 4))
    ref = _oracle_episode(env.spec, "uniform", streams, acts, rule="d+a")
    o1, o2 = env.reset(), env2.reset()
    assert torch.equal(o1, o2)
    for t in range(T - 1):
        o1, r1, _, _ = env.step(torch.as_tensor(acts[t]))
        o2, r2, _, _ = env2.step(torch.as_tensor(acts[t]))
        assert np.array_equal(o1.cpu().numpy(), ref["obs"][t]) and torch.equal(o1, o2) and torch.equal(r1, r2)


def test_all_registry_ids_construct_and_step():
    from multi_echelon_drl_b200 import register_envs as R
    for env_id in R.registered_ids():
        env = R.make(env_id, n_envs=5, seed=1)
        obs = env.reset()
        full = "FullInfo" in env_id
        multi = "MultiFacility" in env_id
        assert obs.shape == (5, 28 if (full or multi) else 7), env_id
        na = 4 if multi else 1
        a = torch.full((5, na), 4.0, device="cuda")
        obs, rew, dones, infos = env.step(a)
        assert obs.shape[0] == 5 and rew.shape == (5,) and dones.shape == (5,) and len(infos) == 5
        assert torch.isfinite(obs).all() and (rew <= 0).all()
    # the one genuinely Discrete id (register_envs.py:27-33); the others keep Box (Q3)
    assert type(R.make("BeerGameUniformRetailerDiscreteDPlusA-v0", n_envs=2).action_space).__name__ == "Discrete"
    assert type(R.make("BeerGameUniformRetailerDiscrete-v0", n_envs=2).action_space).__name__ == "Box"


def test_device_episode_source_distributions_and_determinism():
    from multi_echelon_drl_b200 import register_envs as R
    e1 = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=4096, seed=3)
    e2 = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=4096, seed=3)
    assert torch.equal(e1.reset(), e2.reset())
    d = e1.sim.demand[:, :4096]
    assert d.min() == 0 and d.max() == 8 and abs(d.float().mean().item() - 4.0) < 0.05
    init = e1.sim.init[:, :4096]
    assert init[:4].min() == 0 and init[:4].max() == 24 and init[4:].max() == 8
    e3 = R.make("BeerGameNormalRetailer-v0", n_envs=4096, seed=3)
    e3.reset()
    d = e3.sim.demand[:, :4096]
    assert abs(d.float().mean().item() - 10.0) < 0.1 and abs(d.float().std().item() - 2.0) < 0.1 and (d == d.round()).all()


def test_base_stock_policy_class_on_device_and_hge_contract():
    """HgeTD3.predict's heuristic branch (hge.py:157-167) calls heuristic.get_order_quantity(original_obs):
    per-env actions on device; driving the env with them equals the in-kernel policy rollout."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    idx = {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1, "latest_demand": 2}
    bsp = BaseStockPolicy([19, 20, 20, 14], idx, 7, 0, 16, "a")
    n = 300
    env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=11, episode_source="numpy_bulk")
    obs = env.reset()
    snap = env.sim.get_state()
    total = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(100):
        a = bsp.get_order_quantity(env.get_original_obs())
        assert a.shape == (n, 4) and a.is_cuda
        if t < 99:
            obs, rew, _, _ = env.step(a)
            total += rew.double()
    env.sim.set_state(snap)
    env.sim.rollout(99, policy=BasePolicyParams([19, 20, 20, 14], 0, 16, "a"))
    assert torch.equal(env.sim.ep_return, total)
    out = bsp.get_order_quantity(obs.cpu().numpy())          # numpy in -> numpy float64 out, like the reference
    assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (n, 4)


def test_float_vecenv_accepts_fractional_actions_and_int_rejects_them():
    from multi_echelon_drl_b200 import register_envs as R
    envf = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=8, dtype="float32", seed=1)
    envf.reset()
    obs, rew, _, _ = envf.step(torch.full((8, 4), 3.25, device="cuda"))
    assert (obs[:, 3] == 3.25).all()                     # newest pipeline slot of the retailer = its order
    envi = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=8, dtype="int32", seed=1)
    envi.reset()
    with pytest.raises(TypeError):
        envi.step(torch.full((8, 4), 3.25, device="cuda"))
    # negative orders are clamped to 0 like network_components.py:368-374 (no warning spam on device)
    obs, _, _, _ = envi.step(torch.full((8, 4), -5, device="cuda"))
    assert (obs[:, 3] == 0).all()


def test_multi_item_network_equals_independent_item_chains():
    """Config 4 semantics (parity unpinned: the reference has no multi-item behaviour): two items from
    network_configs/multi_item.yaml = two independent chains with their own holding costs; obs is the
    concatenation, reward the sum.  Checked against the oracle run once per item."""
    import os
    import multi_echelon_drl_b200 as pkg
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    from multi_echelon_drl_b200.spec import specs_from_config
    from multi_echelon_drl_b200.utils.graph import read_yaml
    from oracle import beergame_np as O
    from helpers import split_init
    net = read_yaml(os.path.join(os.path.dirname(pkg.__file__), "network_configs", "multi_item.yaml"))
    ag = ["retailer", "wholesaler", "distributor", "manufacturer"]
    specs = specs_from_config(net, ag, True, 100)
    n, T = 517, 100
    rs = np.random.RandomState(3)
    init = np.concatenate([rs.randint(0, 25, (n, 2, 4)), rs.randint(0, 9, (n, 2, 15))], axis=2)
    demand = rs.randint(0, 9, (n, 2, T + 3))
    acts = rs.randint(0, 17, (T, n, 8))
    sim = MultiItemSupplyNetwork(specs, n, "cuda", "int32")
    assert sim.obs_dim == 56 and sim.n_agents == 8
    sim.load_episode(init, demand)
    obs = sim.reset()
    orcs = []
    for i, h in enumerate(([0.5] * 4, [0.75] * 4)):
        sp = O.OracleSpec(h, [1.0] * 4, [19, 20, 20, 14], (0, 1, 2, 3), True, T)
        orc = O.BeerGameOracle(sp, n)
        inv0, so, sh = split_init(init[:, i])
        o0 = orc.reset(inv0, so, sh, demand[:, i])
        assert np.array_equal(obs[:, 28 * i:28 * (i + 1)].cpu().numpy(), o0)
        orcs.append(orc)
    for t in range(T):
        obs, rew, done = sim.step(torch.as_tensor(acts[t]))
        want_r = np.zeros(n)
        for i, orc in enumerate(orcs):
            o, r, _ = orc.step(acts[t][:, 4 * i:4 * (i + 1)])
            assert np.array_equal(obs[:, 28 * i:28 * (i + 1)].cpu().numpy(), o), (t, i)
            want_r += r
        assert np.array_equal(rew.cpu().numpy(), want_r.astype(np.float32))
    assert done and sim.terminal


def test_run_policy_episodes_statistics_match_oracle():
    """Config-3 path: whole episodes with the in-kernel base-stock policy; statistics vector vs oracle."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.distributed import summarize
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    n, seed, E = 200, 42, 3
    env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=seed, episode_source="numpy_seeded")
    st = env.run_policy_episodes(BaseStockPolicy([19, 20, 20, 14], lb=0, ub=16, rule="a"), E).cpu().numpy()
    streams = [np.random.RandomState(seed + i) for i in range(n)]
    rets, hold, back = [], np.zeros(4), np.zeros(4)
    for _ in range(E):
        ref = _oracle_episode_policy(env.spec, streams)
        rets.append(ref["rew"].sum(0)); hold += ref["holding"].sum((0, 1)); back += ref["backorder"].sum((0, 1))
    rets = np.concatenate(rets)
    assert st[0] == n * 100 * E and st[1] == rets.sum() and st[10] == n * E and st[11] == rets.sum()
    assert np.array_equal(st[2:6], hold) and np.array_equal(st[6:10], back)
    assert abs(st[12] - (rets ** 2).sum()) < 1e-6 * (rets ** 2).sum()
    s = summarize(torch.as_tensor(st))
    assert abs(s["mean_episode_return"] - rets.mean()) < 1e-9 and abs(s["std_episode_return"] - rets.std()) < 1e-6
    assert -1500 < s["mean_episode_return"] < -1100          # SURVEY App. A anchor: about -1300 for this heuristic


def _oracle_episode_policy(spec, streams):
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs_from_streams
    init, demand = seeded_episode_inputs_from_streams(spec, streams)
    return oracle_rollout("uniform", spec.agents, spec.global_observable, "a", 100, init, demand, None,
                          policy=([19, 20, 20, 14], 0, 16, "a"))


def test_kernels_are_cuda_graph_capturable():
    """The library launches on the caller's stream and allocates nothing, so whole episodes of per-period
    calls can be captured in a CUDA graph and replayed (how launch-bound small batches are driven)."""
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.simulator import BasePolicyParams, BatchedSupplyNetwork, heuristic_actions
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    n, T = 1024, 100
    spec = uniform_spec(None, True, T)
    init, demand = bulk_episode_inputs(spec, n, np.random.RandomState(0))
    pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
    sim = BatchedSupplyNetwork(spec, n, "cuda", "int32")
    sim.load_episode(init, demand)
    # eager reference
    obs = sim.reset()
    for _ in range(T):
        obs, _, _ = sim.step(heuristic_actions(pol, obs).to(torch.int32))
    want_ret, want_obs = sim.ep_return.clone(), obs.clone()
    # captured: reset + 100 x (heuristic kernel, cast, step kernel) in one graph
    act_f = torch.zeros((n, 4), device="cuda")
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o = sim.reset()
        for _ in range(T):
            heuristic_actions(pol, o, act_f)
            o, _, _ = sim.step(act_f.to(torch.int32))
    for _ in range(2):                      # replays are idempotent: every replay starts with the reset
        sim.ep_return.fill_(7.0)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(sim.ep_return, want_ret) and torch.equal(sim.obs, want_obs)


def test_philox_table_generation_kernels():
    """sn_fill_uniform / sn_fill_normal_demand: bounds, moments, determinism in (seed, offset), independence of
    the launch geometry (a table equals the concatenation of its parts when offsets continue), ragged counts."""
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = 1_000_003
    for code, dt in ((_lib.SN_I32, torch.int32), (_lib.SN_F32, torch.float32), (_lib.SN_F64, torch.float64)):
        a = torch.full((n + 5,), -7, dtype=dt, device="cuda")
        _lib.check(lib.sn_fill_uniform(code, C.c_void_p(a.data_ptr()), n, 0, 9, 123, 0, st))
        v = a[:n].double()
        assert (a[n:] == -7).all() and v.min() == 0 and v.max() == 8 and (v == v.round()).all()
        assert abs(v.mean().item() - 4.0) < 0.01 and abs(v.var().item() - 80 / 12) < 0.05
        b = torch.empty_like(a)
        _lib.check(lib.sn_fill_uniform(code, C.c_void_p(b.data_ptr()), n, 0, 9, 123, 0, st))
        assert torch.equal(a[:n], b[:n])
        _lib.check(lib.sn_fill_uniform(code, C.c_void_p(b.data_ptr()), n, 0, 9, 124, 0, st))
        assert not torch.equal(a[:n], b[:n])
    # offsets continue the stream: two halves == one table
    whole = torch.empty(4096, dtype=torch.int32, device="cuda")
    parts = torch.empty(4096, dtype=torch.int32, device="cuda")
    _lib.check(lib.sn_fill_uniform(_lib.SN_I32, C.c_void_p(whole.data_ptr()), 4096, 3, 25, 9, 100, st))
    _lib.check(lib.sn_fill_uniform(_lib.SN_I32, C.c_void_p(parts.data_ptr()), 2048, 3, 25, 9, 100, st))
    _lib.check(lib.sn_fill_uniform(_lib.SN_I32, C.c_void_p(parts[2048:].data_ptr()), 2048, 3, 25, 9, 100 + 512, st))
    assert torch.equal(whole, parts) and whole.min() == 3 and whole.max() == 24
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib.sn_fill_normal_demand(_lib.SN_I32, C.c_void_p(d.data_ptr()), n, 10.0, 4.0, 5, 0, st))
    dd = d.double()                                   # the reference's own test: N(10,4) clipped at 0 (tests/test_demands.py:46-49)
    assert abs(dd.mean().item() - 10) < 0.1 and abs(dd.std().item() - 4) < 0.04 and dd.min() == 0
    assert lib.sn_fill_uniform(_lib.SN_I32, C.c_void_p(d.data_ptr()), n, 5, 5, 0, 0, st) == _lib.SN_ERR_INVALID


def test_heuristic_dict_path_matches_reference_golden():
    from conftest import GOLDEN
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    z = np.load(f"{GOLDEN}/heuristic.npz")
    keys = ["on_hand", "unfilled_demand", "latest_demand"] + [f"unreceived_pipeline_{i}" for i in range(4)]
    obs = z["uniform_obs"]
    for rule in ("a", "d+a"):
        for f, tg in enumerate([19, 20, 20, 14]):
            pol = BaseStockPolicy([tg], rule=rule)
            for i in (0, 5, 40, 63):
                st = {k: obs[i, 7 * f + j] for j, k in enumerate(keys)}
                got = pol.get_order_quantity({"x": st})
                assert got.shape == (1, 1)
                np.testing.assert_allclose(got.item(), z[f"uniform_{rule}_dict"][i, f], rtol=1e-6, atol=1e-6)
    both = BaseStockPolicy([19, 14], rule="a").get_order_quantity({"x": {k: obs[0, j] for j, k in enumerate(keys)}})
    st0 = {"x": {k: obs[0, j] for j, k in enumerate(keys)}}
    assert both.shape == (1, 2)                                          # two targets applied to one state
    assert both[0, 0] == BaseStockPolicy([19], rule="a").get_order_quantity(st0).item()
    assert both[0, 1] == BaseStockPolicy([14], rule="a").get_order_quantity(st0).item()
    with pytest.raises(TypeError):
        BaseStockPolicy([1]).get_order_quantity(3)
    with pytest.raises(NotImplementedError):
        BaseStockPolicy([1], array_index={"on_hand": 1, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 0,
                                          "latest_demand": 2})




def test_policy_episodes_replayed_as_graphs_are_the_same_episodes():
    """run_policy_episodes(use_graph=True) replays pairs of episodes as one captured graph that draws the next
    episode's tables in a parallel branch, with the Philox stream position in device memory: same counters in the
    same order, so the statistics are bit-identical to the serial loop - also across calls, odd counts, and a
    reset() / step() in between."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
    out = []
    for use_graph in (False, True):
        env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=20000, dtype="int32", seed=5)
        st = env.run_policy_episodes(pol, 7, use_graph=use_graph).clone()
        st2 = env.run_policy_episodes(pol, 4, use_graph=use_graph).clone()    # a second call continues the streams
        obs = env.reset().clone()                                             # ... and so does an ordinary reset
        st3 = env.run_policy_episodes(pol, 3, use_graph=use_graph).clone()
        out.append((st, st2, obs, st3))
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)
    assert out[0][0][0].item() == 20000 * 100 * 7 and out[0][0][10].item() == 20000 * 7
