"""GPU parity tests: CUDA path (through the C ABI) vs golden vectors of the unmodified reference
and vs the numpy oracle on seeded inputs.  Integer quantities must be bit-exact; float paths are
held to 1e-6 relative to the scale of the state (tolerances written at each assert)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import CLASSIC_TOTALS, load_golden_classic, load_golden_rollouts  # noqa: E402
from helpers import DPA, oracle_rollout  # noqa: E402


def _make(scenario, agents, glob, rule, T, n, dtype):
    from multi_echelon_drl_b200 import classic_spec, normal_spec, uniform_spec
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    if scenario == "uniform":
        spec = uniform_spec(agents, glob, T)
    elif scenario == "normal":
        spec = normal_spec(agents, glob, T)
    else:
        spec = classic_spec(agents, T)
    if rule == "d+a":
        spec = spec.with_rule("d+a", *DPA[scenario])
    return BatchedSupplyNetwork(spec, n, "cuda", dtype)


GOLD = load_golden_rollouts()


@pytest.mark.parametrize("case", GOLD, ids=[c["id"] for c in GOLD])
def test_step_matches_reference_golden(case):
    """sn_reset + sn_step, one env, against trajectories recorded from the reference itself."""
    is_int = case["kind"] == "int"
    sim = _make(case["scenario"], case["agents"], case["global_observable"], case["rule"], 100, 1,
                "int32" if is_int else "float32")
    sim.load_episode(case["init"][None], case["demand"][None])
    obs0 = sim.reset().cpu().numpy()
    T = case["actions"].shape[0]
    cost = torch.zeros((8, sim.ld), dtype=torch.float64, device="cuda")
    r64 = torch.zeros((1,), dtype=torch.float64, device="cuda")
    # float cases: BASELINE.json's tolerance, 1e-6 relative - to the scale of the trajectory (max |obs|, max |reward|),
    # the only meaningful reference for a state in which 0.0 is a legitimate value.  The reference itself computes
    # these rollouts in float32 scalars (numpy >= 2); measured worst case over the 24 float goldens on a B200:
    # 4.5e-7 (obs), 6.9e-7 (reward) - profiles/float_error_probe.py
    oscale, rscale = max(1.0, float(np.abs(case["obs"]).max())), max(1.0, float(np.abs(case["rew"]).max()))
    if is_int:
        assert np.array_equal(obs0[0], case["obs0"])
    else:
        assert np.abs(obs0[0] - case["obs0"]).max() <= 1e-6 * oscale
    for t in range(T):
        obs, rew, done = sim.step(torch.as_tensor(case["actions"][t][None]), reward64_out=r64, cost_out=cost)
        o, r, c = obs.cpu().numpy()[0], r64.item(), cost[:, 0].cpu().numpy()
        if is_int:      # bit-exact: integer quantities, dyadic cost coefficients
            assert np.array_equal(o, case["obs"][t]), t
            assert r == case["rew"][t], t
            assert np.array_equal(c[:4], case["holding"][t]) and np.array_equal(c[4:], case["backorder"][t]), t
            assert rew.item() == np.float32(case["rew"][t])
        else:           # float32 arithmetic vs the reference's mixed float32/python arithmetic
            assert np.abs(o.astype(np.float64) - case["obs"][t]).max() <= 1e-6 * oscale, t
            assert abs(r - case["rew"][t]) <= 1e-6 * rscale, t
        assert done == bool(case["done"][t])
    with pytest.raises(ValueError):
        sim.step(torch.as_tensor(case["actions"][0][None]))            # envs.py:88-89


@pytest.mark.parametrize("case", [c for c in GOLD if c["kind"] == "int"][::3], ids=lambda c: c["id"])
def test_rollout_matches_reference_golden(case):
    """sn_rollout (fused, streamed actions) reproduces the same golden trajectory, incl. terminal obs."""
    sim = _make(case["scenario"], case["agents"], case["global_observable"], case["rule"], 100, 1, "int32")
    sim.load_episode(case["init"][None], case["demand"][None])
    sim.reset()
    acts = torch.as_tensor(case["actions"][:, None, :])
    out = sim.rollout(40, actions=acts[:40], record_obs=True, record_reward=True)     # split: state write-back
    out2 = sim.rollout(60, actions=acts[40:], record_obs=True, record_reward=True)
    obs = torch.cat([out["traj_obs"], out2["traj_obs"]]).cpu().numpy()[:, 0]
    rew = torch.cat([out["traj_reward"], out2["traj_reward"]]).cpu().numpy()[:, 0]
    assert np.array_equal(obs, case["obs"])
    assert np.array_equal(rew, case["rew"].astype(np.float32))
    assert np.array_equal(out2["obs"].cpu().numpy()[0], case["obs"][-1]) and out2["done"]
    assert sim.ep_return.item() == case["rew"].sum()


def test_classic_beer_game_totals():
    """The 13 golden episode totals of the reference's tests/test_supplynetwork.py:25-69."""
    for ep, total in zip(load_golden_classic(), CLASSIC_TOTALS):
        sim = _make("classic", tuple(int(a) for a in ep["agents"]), False, "a", 35, 1, "int32")
        sim.load_episode(ep["init"][None], ep["demand"][None])
        assert np.array_equal(sim.reset().cpu().numpy()[0], ep["obs0"])
        tot = 0.0
        for t in range(20):
            obs, rew, done = sim.step(torch.as_tensor(ep["actions"][t][None]))
            assert np.array_equal(obs.cpu().numpy()[0], ep["obs"][t])
            tot += rew.item()
        assert tot == total


def _seeded_batch(scenario, n, T, seed):
    from multi_echelon_drl_b200 import normal_spec, uniform_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    spec = uniform_spec(max_episode_steps=T) if scenario == "uniform" else normal_spec(max_episode_steps=T, random_init=True)
    return bulk_episode_inputs(spec, n, np.random.RandomState(seed))


@pytest.mark.parametrize("scenario,agents,glob,rule", [
    ("uniform", (0, 1, 2, 3), True, "a"), ("uniform", (0, 1, 2, 3), False, "d+a"),
    ("normal", (0,), False, "a"), ("normal", (1, 2), True, "a"), ("uniform", (3,), True, "d+a"),
    ("uniform", (3, 2, 1, 0), False, "a"),
])
def test_batch_step_vs_oracle(scenario, agents, glob, rule):
    """4,133 envs (ragged: not a multiple of 32/128) x 100 steps, step kernel vs numpy oracle, bit-exact."""
    n, T = 4133, 100
    init, demand = _seeded_batch(scenario, n, T, 11)
    hi = 17 if scenario == "uniform" else 21
    acts = np.random.RandomState(5).randint(-1, hi + 2, (T, n, len(agents)))
    ref = oracle_rollout(scenario, agents, glob, rule, T, init, demand, acts)
    sim = _make(scenario, agents, glob, rule, T, n, "int32")
    sim.track_stats = True
    sim.load_episode(init, demand)
    assert np.array_equal(sim.reset().cpu().numpy(), ref["obs0"])
    r64 = torch.zeros((n,), dtype=torch.float64, device="cuda")
    for t in range(T):
        obs, rew, done = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
        assert np.array_equal(obs.cpu().numpy(), ref["obs"][t]), t
        assert np.array_equal(r64.cpu().numpy(), ref["rew"][t]), t
    assert done
    st = sim.stats.cpu().numpy()
    assert st[0] == n * T and st[1] == ref["rew"].sum()
    assert np.array_equal(st[2:6], ref["holding"].sum((0, 1))) and np.array_equal(st[6:10], ref["backorder"].sum((0, 1)))
    assert np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 129, 1000])
def test_ragged_sizes_rollout(n):
    """Edge sizes around warp / block boundaries: fused rollout with trajectory vs oracle."""
    T = 30
    init, demand = _seeded_batch("uniform", n, T, n)
    acts = np.random.RandomState(n).randint(0, 17, (T, n, 4))
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts)
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
    sim.load_episode(init, demand)
    sim.reset()
    out = sim.rollout(T, actions=torch.as_tensor(acts), record_obs=True, record_reward=True)
    assert np.array_equal(out["traj_obs"].cpu().numpy(), ref["obs"])
    assert np.array_equal(out["traj_reward"].cpu().numpy(), ref["rew"].astype(np.float32))


@pytest.mark.parametrize("prule,erule", [("a", "a"), ("d+a", "d+a"), ("a", "d+a")])
@pytest.mark.parametrize("scenario", ["uniform", "normal"])
def test_heuristic_rollout_vs_oracle(scenario, prule, erule):
    """Config-3 mode: base-stock policy evaluated in-kernel per env (utils/heuristics.py), fused
    100-step rollout, vs the oracle driven by the oracle's own heuristic restatement."""
    n, T = 2050, 100
    init, demand = _seeded_batch(scenario, n, T, 3)
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    tg, ub = ([19, 20, 20, 14], 16) if scenario == "uniform" else ([48, 43, 41, 30], 20)
    ref = oracle_rollout(scenario, (0, 1, 2, 3), True, erule, T, init, demand, None, policy=(tg, 0, ub, prule))
    sim = _make(scenario, (0, 1, 2, 3), True, erule, T, n, "int32")
    sim.track_stats = True
    sim.load_episode(init, demand)
    sim.reset()
    out = sim.rollout(T, policy=BasePolicyParams(tg, 0, ub, prule), record_obs=True, record_reward=True,
                      record_actions=True)
    assert np.array_equal(out["traj_actions"].cpu().numpy(), ref["actions"].astype(np.float32))
    assert np.array_equal(out["traj_obs"].cpu().numpy(), ref["obs"])
    assert np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))
    assert sim.stats[1].item() == ref["rew"].sum()


def test_heuristic_kernel_vs_reference_golden():
    """sn_heuristic vs BaseStockPolicy.get_order_quantity outputs recorded from the reference
    (ndarray path and dict path agree there; tolerance 1e-6 relative for the fractional rows)."""
    import os
    from conftest import GOLDEN
    from multi_echelon_drl_b200.simulator import BasePolicyParams, heuristic_actions
    z = np.load(os.path.join(GOLDEN, "heuristic.npz"))
    for name, tg, ub in (("uniform", [19, 20, 20, 14], 16), ("normal", [48, 43, 41, 30], 20)):
        obs = torch.as_tensor(z[f"{name}_obs"]).cuda()
        for rule in ("a", "d+a"):
            got = heuristic_actions(BasePolicyParams(tg, 0, ub, rule), obs).cpu().numpy()
            want = z[f"{name}_{rule}_nd"][:, 0, :] if z[f"{name}_{rule}_nd"].ndim == 3 else z[f"{name}_{rule}_nd"]
            assert np.array_equal(got[:32], want[:32].astype(np.float32))          # integer rows: exact
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(got, z[f"{name}_{rule}_dict"], rtol=1e-6, atol=1e-5)
        # the reference's batch quirk: only row 0 is read
        got = heuristic_actions(BasePolicyParams(tg, 0, ub, "a", row0_broadcast=True), obs[:2]).cpu().numpy()
        np.testing.assert_allclose(got[0], z[f"{name}_row0_batch2"].reshape(-1), rtol=1e-6)
        np.testing.assert_allclose(got[1], z[f"{name}_row0_batch2"].reshape(-1), rtol=1e-6)


def test_float_paths_vs_oracle():
    """Fractional (TD3-like) orders: f32 and f64 kernels vs the float64 oracle, BASELINE.json's 1e-6 relative (to the
    scale of the trajectory).  f64: the float32 outputs are the correctly rounded oracle values (1e-7 covers the
    rounding of the output itself; measured 0 / 4.6e-8); f32: 1e-6 (measured 4.7e-7 obs, 3.1e-7 reward after 100
    compounding periods, profiles/float_error_probe.py)."""
    n, T = 1500, 100
    init, demand = _seeded_batch("uniform", n, T, 9)
    acts = (np.random.RandomState(2).rand(T, n, 4) * 16).astype(np.float32)
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts.astype(np.float64), dtype=np.float64)
    scale = max(1.0, np.abs(ref["obs"]).max())
    for dt, tol in (("float64", 1e-7), ("float32", 1e-6)):
        sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, dt)
        sim.load_episode(init, demand)
        sim.reset()
        out = sim.rollout(T, actions=torch.as_tensor(acts), record_obs=True, record_reward=True)
        err = np.abs(out["traj_obs"].cpu().numpy().astype(np.float64) - ref["obs"]).max()
        assert err <= tol * scale, (dt, err, scale)
        rerr = np.abs(out["traj_reward"].cpu().numpy() - ref["rew"]).max()
        assert rerr <= tol * np.abs(ref["rew"]).max(), (dt, rerr)


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (65,536 envs, complex scenario, centralized, fixed actions):
    (i) step-by-step kernel and fused rollout agree bit-for-bit; (ii) per-env results are invariant
    to batch position (first 4,096 envs equal a 4,096-env run = shard invariance); (iii) the
    reference's runtime invariant on_order == sum(pipeline) (supplynetwork.py:159-160) holds."""
    n, T = 65536, 100
    init, demand = _seeded_batch("uniform", n, T, 1234)
    acts = np.random.RandomState(77).randint(0, 17, (T, n, 4)).astype(np.int32)
    a = torch.as_tensor(acts).cuda()
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
    sim.load_episode(init, demand)
    sim.reset()
    out = sim.rollout(T, actions=a, record_obs=True, record_reward=True)
    ret_fused = sim.ep_return.clone()
    sim2 = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
    sim2.load_episode(init, demand)
    sim2.reset()
    for t in range(T):
        obs, rew, _ = sim2.step(a[t])
        assert torch.equal(obs, out["traj_obs"][t]) and torch.equal(rew, out["traj_reward"][t])
        if t == 50:
            s = sim2.state[:, :n].to(torch.int64)
            for k in range(4):
                on_order = s[7 * k + 3:7 * k + 7].sum(0)
                pipe = s[31 + 2 * k] + s[32 + 2 * k] + (s[28 + k] if k < 3 else 0) + (s[7 * (k + 1) + 1] if k < 3 else s[39])
                assert torch.equal(on_order, pipe)
    assert torch.equal(ret_fused, sim2.ep_return)
    m = 4096
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init[:m], demand[:m], acts[:, :m])
    assert np.array_equal(out["traj_obs"][:, :m].cpu().numpy(), ref["obs"])
    assert np.array_equal(ret_fused[:m].cpu().numpy(), ref["rew"].sum(0))
    # (iv) ALL 65,536 envs x 100 periods against the C restatement of the reference (itself pinned to the goldens and
    # to the numpy restatement by tests/test_oracle_c.py): observations, rewards, episode returns, bit-exact
    import os
    from oracle import beergame_c as OC
    from oracle import beergame_np as O
    full = OC.rollout(O.scenario_spec("uniform", (0, 1, 2, 3), True, T), init, demand, acts, n_threads=os.cpu_count() or 1)
    assert np.array_equal(out["traj_obs"].cpu().numpy(), full["obs"])
    assert np.array_equal(out["traj_reward"].cpu().numpy(), full["rew"].astype(np.float32))
    assert np.array_equal(ret_fused.cpu().numpy(), full["rew"].sum(0))


@pytest.mark.parametrize("scenario,agents,glob,rule", [
    ("uniform", (0, 1, 2, 3), True, "a"), ("normal", (1, 2, 3), False, "d+a"), ("uniform", (2,), True, "a"),
])
def test_episode_fused_reset_rollout_vs_oracle(scenario, agents, glob, rule):
    """sn_episode: in-kernel reset + rollout in one launch (state never in HBM) == reset + steps."""
    n, T = 3001, 100
    init, demand = _seeded_batch(scenario, n, T, 17)
    hi = 17 if scenario == "uniform" else 21
    acts = np.random.RandomState(8).randint(0, hi, (T, n, len(agents)))
    ref = oracle_rollout(scenario, agents, glob, rule, T, init, demand, acts)
    sim = _make(scenario, agents, glob, rule, T, n, "int32")
    sim.load_episode(init, demand)
    obs0 = torch.zeros((n, sim.obs_dim), device="cuda")
    sim.ep_return.fill_(123.0)                                  # must be overwritten, not accumulated
    out = sim.episode(T, actions=torch.as_tensor(acts), record_obs=True, record_reward=True, obs0_out=obs0)
    assert np.array_equal(obs0.cpu().numpy(), ref["obs0"])
    assert np.array_equal(out["traj_obs"].cpu().numpy(), ref["obs"])
    assert np.array_equal(out["traj_reward"].cpu().numpy(), ref["rew"].astype(np.float32))
    assert np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))
    assert np.array_equal(out["obs"].cpu().numpy(), ref["obs"][-1]) and out["done"]
    # partial episode with state write-back, continued by single steps
    out = sim.episode(60, actions=torch.as_tensor(acts[:60]))
    assert not out["done"] and sim.period == 60
    for t in range(60, T):
        obs, _, done = sim.step(torch.as_tensor(acts[t]))
        assert np.array_equal(obs.cpu().numpy(), ref["obs"][t])
    assert done and np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))


@pytest.mark.parametrize("chunks", [1, 4, 7])
def test_episode_host_pipelined_equals_plain(chunks):
    """The pipelined host-buffer path (H2D / kernels / D2H overlapped over spans of periods) returns
    exactly the trajectory of the plain path."""
    n, T = 5000, 100
    init, demand = _seeded_batch("uniform", n, T, 23)
    acts = np.random.RandomState(4).randint(0, 17, (T, n, 4)).astype(np.int32)
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts)
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
    h_obs = torch.empty((T, n, 28), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((T, n), dtype=torch.float32).pin_memory()
    h_act = torch.from_numpy(acts).pin_memory()
    for _ in range(2):                                      # twice: persistent pipeline buffers are reused
        h_obs.zero_(); h_rew.zero_()
        sim.episode_host(torch.from_numpy(init), torch.from_numpy(demand), h_act, h_obs, h_rew, chunks=chunks)
        assert np.array_equal(h_obs.numpy(), ref["obs"])
        assert np.array_equal(h_rew.numpy(), ref["rew"].astype(np.float32))
        assert np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))


@pytest.mark.parametrize("n", [1, 33, 95, 130])
def test_no_out_of_bounds_writes_canaries(n):
    """compute-sanitizer is closed on this pool, so bounds are checked with guard bands: every output
    buffer is a window inside a larger sentinel-filled allocation; padding lanes of the state rows
    ([n, ld)) must stay untouched as well."""
    T, G = 12, 64
    init, demand = _seeded_batch("uniform", n, T, 5)
    acts = np.random.RandomState(1).randint(0, 17, (T, n, 4))
    for agents, glob in (((0, 1, 2, 3), True), ((1,), False), ((2, 1), True)):
        sim = _make("uniform", agents, glob, "a", T, n, "int32")
        od, na = sim.obs_dim, sim.n_agents
        sim.state.fill_(-77)
        sim.load_episode(init, demand)

        def window(shape, dtype=torch.float32):
            numel = int(np.prod(shape))
            big = torch.full((numel + 2 * G,), -12345.0 if dtype.is_floating_point else -12345, dtype=dtype, device="cuda")
            return big, big[G:G + numel].view(shape)

        b_obs, obs = window((n, od))
        b_to, traj_obs = window((T, n, od))
        b_tr, traj_rew = window((T, n))
        b_ta, traj_act = window((T, n, na))
        b_rw, rew = window((n,))
        sim.reset(obs_out=obs)
        sim.step(torch.as_tensor(acts[0][:, :na]), obs_out=obs, reward_out=rew)
        sim.rollout(T - 1, actions=torch.as_tensor(acts[1:, :, :na]), traj_obs=traj_obs[1:], traj_reward=traj_rew[1:],
                    traj_actions=traj_act[1:])
        sim.episode(T, actions=torch.as_tensor(acts[:, :, :na]), traj_obs=traj_obs, traj_reward=traj_rew,
                    traj_actions=traj_act, obs0_out=obs)
        torch.cuda.synchronize()
        for big in (b_obs, b_to, b_tr, b_ta, b_rw):
            assert (big[:G] == -12345).all() and (big[-G:] == -12345).all()
        assert (sim.state[:, n:] == -77).all()
        assert torch.isfinite(traj_obs).all() and (traj_obs != -12345).all()


def test_unaligned_output_buffers_take_the_plain_store_path():
    """Observation / trajectory pointers that are not 16-byte aligned cannot use TMA bulk stores; the
    kernels must fall back to plain stores and still be exact.  Unaligned ACTION rows are rejected."""
    n, T = 257, 10
    init, demand = _seeded_batch("uniform", n, T, 6)
    acts = np.random.RandomState(2).randint(0, 17, (T, n, 4))
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts)
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
    sim.load_episode(init, demand)
    big = torch.zeros(T * n * 28 + 8, device="cuda")
    traj = big[1:1 + T * n * 28].view(T, n, 28)                    # 4-byte offset: misaligned for TMA
    assert traj.data_ptr() % 16 != 0
    obs0 = torch.zeros(n * 28 + 8, device="cuda")[3:3 + n * 28].view(n, 28)
    sim.episode(T, actions=torch.as_tensor(acts), traj_obs=traj, obs0_out=obs0)
    assert np.array_equal(traj.cpu().numpy(), ref["obs"]) and np.array_equal(obs0.cpu().numpy(), ref["obs0"])
    o = torch.zeros(n * 28 + 8, device="cuda")[1:1 + n * 28].view(n, 28)
    sim.reset(obs_out=o)
    assert np.array_equal(o.cpu().numpy(), ref["obs0"])
    sim.step(torch.as_tensor(acts[0]), obs_out=o)
    assert np.array_equal(o.cpu().numpy(), ref["obs"][0])
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    a = torch.zeros(n * 4 + 4, dtype=torch.int32, device="cuda")[1:]
    rc = sim.lib.sn_step(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, 1, C.c_void_p(sim.state.data_ptr()),
                         C.c_void_p(a.data_ptr()), C.c_void_p(sim.demand[2].data_ptr()), None, None, None, None, None,
                         None, None)
    assert rc == _lib.SN_ERR_INVALID and b"aligned" in sim.lib.sn_last_error()


def test_full_size_config3_heuristic_paths_agree():
    """BASELINE config 3, one GPU's share at full size (131,072 envs, base-stock [19,20,20,14], rule 'a'):
    (i) the in-kernel policy rollout (sn_episode) equals driving sn_step with sn_heuristic's actions
    computed from the observations; (ii) d+a heuristic through the d+a env rule likewise; (iii) a
    4,096-env prefix equals the oracle; (iv) mean return near the reference anchor (-1300)."""
    from multi_echelon_drl_b200.simulator import BasePolicyParams, heuristic_actions
    n, T = 131072, 100
    init, demand = _seeded_batch("uniform", n, T, 31)
    for rule in ("a", "d+a"):
        pol = BasePolicyParams([19, 20, 20, 14], 0, 16, rule)
        sim = _make("uniform", (0, 1, 2, 3), True, rule, T, n, "int32")
        sim.load_episode(init, demand)
        sim.episode(T, policy=pol, write_state=False)
        fused = sim.ep_return.clone()
        obs = sim.reset()
        for t in range(T):
            obs, _, _ = sim.step(heuristic_actions(pol, obs))
        assert torch.equal(fused, sim.ep_return), rule
        m = 4096
        ref = oracle_rollout("uniform", (0, 1, 2, 3), True, rule, T, init[:m], demand[:m], None,
                             policy=([19, 20, 20, 14], 0, 16, rule))
        assert np.array_equal(fused[:m].cpu().numpy(), ref["rew"].sum(0)), rule
        if rule == "a":
            assert -1400 < fused.mean().item() < -1200


def test_full_size_config4_multi_item_properties():
    """BASELINE config 4 at full size (262,144 envs x 2 items = 524,288 interleaved chains): (i) item 0 of the
    multi-item batch equals a single-item run on the same streams, bit for bit; (ii) item 1 differs only in the
    holding cost (rewards scale, observations identical to a run with its own cost vector); (iii) a prefix equals
    the oracle run per item."""
    import os
    import multi_echelon_drl_b200 as pkg
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    from multi_echelon_drl_b200.spec import specs_from_config
    from multi_echelon_drl_b200.utils.graph import read_yaml
    n, T = 262144, 25
    net = read_yaml(os.path.join(os.path.dirname(pkg.__file__), "network_configs", "multi_item.yaml"))
    specs = specs_from_config(net, ["retailer", "wholesaler", "distributor", "manufacturer"], True, T)
    rs = np.random.RandomState(8)
    init = np.concatenate([rs.randint(0, 25, (n, 2, 4)), rs.randint(0, 9, (n, 2, 15))], axis=2).astype(np.int32)
    demand = rs.randint(0, 9, (n, 2, T + 3)).astype(np.int32)
    acts = torch.randint(0, 17, (T, n, 8), device="cuda", dtype=torch.int32)
    mi = MultiItemSupplyNetwork(specs, n, "cuda", "int32")
    mi.load_episode(init, demand)
    mi.reset()
    singles = []
    for i in range(2):
        s = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "int32")
        if i == 1:
            from dataclasses import replace
            from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
            s = BatchedSupplyNetwork(replace(s.spec, holding_cost=(0.75,) * 4), n, "cuda", "int32")
        s.load_episode(init[:, i], demand[:, i])
        s.reset()
        singles.append(s)
    for t in range(T):
        obs, rew, done = mi.step(acts[t])
        tot = torch.zeros(n, dtype=torch.float64, device="cuda")
        for i, s in enumerate(singles):
            r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
            o, _, _ = s.step(acts[t][:, 4 * i:4 * i + 4].contiguous(), reward64_out=r64)
            assert torch.equal(obs[:, 28 * i:28 * i + 28], o), (t, i)
            tot += r64
        assert torch.equal(rew, tot.float()), t
    assert done
    m = 2048
    for i, h in enumerate(([0.5] * 4, [0.75] * 4)):
        from oracle import beergame_np as O
        from helpers import split_init
        orc = O.BeerGameOracle(O.OracleSpec(h, [1.0] * 4, [19, 20, 20, 14], (0, 1, 2, 3), True, T), m)
        inv0, so, sh = split_init(init[:m, i])
        orc.reset(inv0, so, sh, demand[:m, i])
        ret = np.zeros(m)
        a_np = acts[:, :m, 4 * i:4 * i + 4].cpu().numpy()
        for t in range(T):
            _, r, _ = orc.step(a_np[t])
            ret += r
        assert np.array_equal(singles[i].ep_return[:m].cpu().numpy(), ret)
    assert torch.equal(mi.ep_return, singles[0].ep_return + singles[1].ep_return)


def test_config5_shape_float_actions_vs_oracle():
    """Config-5 shape (4,096 envs, float32 simulator, fractional TD3-like actions through the d+a-free env):
    float32 kernels vs the float64 oracle within 1e-6 of the state scale over a whole episode."""
    n, T = 4096, 100
    init, demand = _seeded_batch("uniform", n, T, 44)
    acts = (np.random.RandomState(4).rand(T, n, 4) * 16).astype(np.float32)
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, acts.astype(np.float64), dtype=np.float64)
    sim = _make("uniform", (0, 1, 2, 3), True, "a", T, n, "float32")
    sim.load_episode(init, demand)
    sim.reset()
    scale = max(1.0, float(np.abs(ref["obs"]).max()))
    for t in range(T):
        obs, rew, _ = sim.step(torch.as_tensor(acts[t]))
        assert np.abs(obs.cpu().numpy() - ref["obs"][t]).max() <= 1e-6 * scale, t
    assert np.abs(sim.ep_return.cpu().numpy() - ref["rew"].sum(0)).max() <= 1e-6 * np.abs(ref["rew"].sum(0)).max()
