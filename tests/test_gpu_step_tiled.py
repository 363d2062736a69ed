"""The TMA-tiled persistent step kernel (k_step_tiled), the device period counter (sn_step_ex ctl), the in-kernel
multi-item reward reduction, the fused VecNormalize statistics and the int32 range guard.

Every case runs the SAME period step twice, once per kernel (SN_STEP_KERNEL=plain / tiled), and demands bit-equal
state, observation, reward, cost, statistics; the plain kernel is itself pinned to the reference goldens and the
oracle by test_gpu_parity.py, and the headline sizes are compared with the oracle directly below."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import oracle_rollout  # noqa: E402


def _spec(scenario, agents, glob, rule, T):
    from multi_echelon_drl_b200 import normal_spec, uniform_spec
    from dataclasses import replace
    names = ("retailer", "wholesaler", "distributor", "manufacturer")
    ag = [names[a] for a in agents]
    sp = uniform_spec(ag, glob, T) if scenario == "uniform" else normal_spec(ag, glob, T, random_init=True)
    if rule == "d+a":
        off = -8 if scenario == "uniform" else -10
        sp = replace(sp, rule="d+a", dpa_offset=off, dpa_lb=0, dpa_ub=sp.action_high)
    return sp


class _Kernel:
    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("SN_STEP_KERNEL")
        os.environ["SN_STEP_KERNEL"] = self.which

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("SN_STEP_KERNEL", None)
        else:
            os.environ["SN_STEP_KERNEL"] = self.old


def _run(which, sp, n, dtype, init, demand, acts, steps, obs_offset=0):
    """`steps` periods through BatchedSupplyNetwork.step with every optional output; returns host copies."""
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    with _Kernel(which):
        sim = BatchedSupplyNetwork(sp, n, "cuda", dtype)
        sim.track_stats = True
        sim.load_episode(init, demand)
        out = dict(obs=[], rew=[], r64=[], cost=[])
        obs_store = torch.zeros((n * sp.obs_dim + 4,), dtype=torch.float32, device="cuda")
        obs_buf = obs_store[obs_offset:obs_offset + n * sp.obs_dim].view(n, sp.obs_dim)   # offset 1: not 16-byte aligned
        out["obs0"] = sim.reset(obs_buf).cpu().numpy().copy()
        r64 = torch.zeros((n,), dtype=torch.float64, device="cuda")
        cost = torch.zeros((8, sim.ld), dtype=torch.float64, device="cuda")
        for t in range(steps):
            obs, rew, done = sim.step(torch.as_tensor(acts[t]), obs_out=obs_buf, reward64_out=r64, cost_out=cost)
            out["obs"].append(obs.cpu().numpy().copy())
            out["rew"].append(rew.cpu().numpy().copy())
            out["r64"].append(r64.cpu().numpy().copy())
            out["cost"].append(cost[:, :n].cpu().numpy().copy())
        out["state"] = sim.state[:, :n].cpu().numpy().copy()
        out["stats"] = sim.stats.cpu().numpy().copy()
        out["ep_return"] = sim.ep_return.cpu().numpy().copy()
        out["done"] = done
    return out


def _same(a, b, exact_stats=True):
    for k in ("obs0", "state", "ep_return"):
        assert np.array_equal(a[k], b[k]), k
    if exact_stats:                       # integer quantities, dyadic cost coefficients: every partial sum is exact
        assert np.array_equal(a["stats"], b["stats"])
    else:                                 # fractional quantities: the block-level summation order differs
        np.testing.assert_allclose(a["stats"], b["stats"], rtol=1e-12)
    for k in ("obs", "rew", "r64", "cost"):
        for t, (x, y) in enumerate(zip(a[k], b[k])):
            assert np.array_equal(x, y), (k, t)
    assert a["done"] == b["done"]


def _inputs(sp, n, T, seed, lo=-1, hi=18):
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    init, demand = bulk_episode_inputs(sp, n, np.random.RandomState(seed))
    acts = np.random.RandomState(seed + 1).randint(lo, hi, (T, n, sp.n_agents))
    return init, demand, acts


@pytest.mark.parametrize("n", [1, 31, 127, 128, 129, 1000, 128 * 148 + 1, 128 * 148 * 3 + 77])
def test_tiled_equals_plain_ragged_sizes(n):
    """One tile per CTA, partial tiles, several tiles per CTA (more tiles than SMs x stages)."""
    T = 12
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 3)
    _same(_run("plain", sp, n, "int32", init, demand, acts, T), _run("tiled", sp, n, "int32", init, demand, acts, T))


@pytest.mark.parametrize("scenario,agents,glob,rule,dtype", [
    ("uniform", (0, 1, 2, 3), False, "d+a", "int32"), ("normal", (0,), False, "a", "int32"),
    ("normal", (1, 2), True, "a", "float32"), ("uniform", (3,), True, "d+a", "float64"),
    ("uniform", (3, 2, 1, 0), False, "a", "float32"), ("uniform", (0, 1, 2, 3), True, "a", "float64"),
    ("uniform", (2, 3), False, "a", "int32"),
])
def test_tiled_equals_plain_layouts_and_dtypes(scenario, agents, glob, rule, dtype):
    """Agent subsets / caller orders, local and global observation, both rules, all three quantity types, whole
    episodes including the terminal step."""
    n, T = 4133, 30
    sp = _spec(scenario, agents, glob, rule, T)
    init, demand, acts = _inputs(sp, n, T, 7)
    if dtype != "int32":
        acts = acts + np.random.RandomState(9).rand(*acts.shape)           # fractional orders
    a, b = _run("plain", sp, n, dtype, init, demand, acts, T), _run("tiled", sp, n, dtype, init, demand, acts, T)
    _same(a, b, exact_stats=(dtype == "int32"))
    assert b["done"]


def test_tiled_unaligned_observation_buffer():
    """An observation buffer that is not 16-byte aligned cannot be bulk-stored: rows are copied out by the threads."""
    n, T = 1000, 8
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 5)
    _same(_run("plain", sp, n, "int32", init, demand, acts, T, obs_offset=1),
          _run("tiled", sp, n, "int32", init, demand, acts, T, obs_offset=1))


def test_tiled_config2_full_size_vs_oracle():
    """BASELINE config 2 at full size through the tiled step kernel: all 65,536 envs x 100 periods against the C
    restatement of the reference (observation, reward, per-facility cost sums), bit-exact."""
    from oracle import beergame_c as OC
    from oracle import beergame_np as O
    n, T = 65536, 100
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 1234, 0, 17)
    ref = OC.rollout(O.scenario_spec("uniform", (0, 1, 2, 3), True, T), init, demand, acts, n_threads=os.cpu_count())
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    with _Kernel("tiled"):
        sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
        sim.load_episode(init, demand)
        assert np.array_equal(sim.reset().cpu().numpy(), ref["obs0"])
        r64 = torch.zeros((n,), dtype=torch.float64, device="cuda")
        d_acts = torch.as_tensor(acts.astype(np.int32)).cuda()
        for t in range(T):
            obs, rew, done = sim.step(d_acts[t], reward64_out=r64)
            assert np.array_equal(obs.cpu().numpy(), ref["obs"][t]), t
            assert np.array_equal(r64.cpu().numpy(), ref["rew"][t]), t
        assert done and np.array_equal(sim.ep_return.cpu().numpy(), ref["rew"].sum(0))


@pytest.mark.parametrize("which", ["plain", "tiled"])
def test_device_period_counter_replays_one_graph(which):
    """sn_step_ex with ctl: ONE captured launch, replayed, walks through the episode exactly like host-period steps;
    replaying past the end does nothing and raises the sticky flag."""
    from multi_echelon_drl_b200 import _lib
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    n, T = 3000, 20
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 21)
    ref = _run(which, sp, n, "int32", init, demand, np.repeat(acts[:1], T, 0), T)
    with _Kernel(which):
        sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
        sim.track_stats = True
        sim.load_episode(init, demand)
        sim.reset()
        a = torch.as_tensor(acts[0].astype(np.int32)).cuda()
        obs = torch.zeros((n, 28), dtype=torch.float32, device="cuda")
        rew = torch.zeros((n,), dtype=torch.float32, device="cuda")
        sim._graph_period_valid = True
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sim.launch_step_device_period(a, obs, rew)
        for t in range(T):
            g.replay()
            assert int(sim.ctl[_lib.SN_CTL_PERIOD].item()) == t + 1
            assert np.array_equal(obs.cpu().numpy(), ref["obs"][t]), t
            assert np.array_equal(rew.cpu().numpy(), ref["rew"][t]), t
        assert np.array_equal(sim.stats.cpu().numpy(), ref["stats"])
        assert np.array_equal(sim.ep_return.cpu().numpy(), ref["ep_return"])
        assert sim.device_flags() == 0
        before = sim.state.clone()
        g.replay()                                                       # period == T: refused on the device
        assert sim.device_flags() & _lib.SN_FLAG_STEP_AFTER_TERMINAL
        assert torch.equal(before, sim.state) and int(sim.ctl[_lib.SN_CTL_PERIOD].item()) == T
        assert int(sim.ctl[_lib.SN_CTL_TICKET].item()) == 0


@pytest.mark.parametrize("n_items,n", [(2, 5000), (4, 777), (2, 128 * 148 + 3)])
@pytest.mark.parametrize("which", ["plain", "tiled"])
def test_multi_item_reward_reduced_in_kernel(which, n_items, n):
    """env_reward = sum over the items' chain rewards, in item order, in double (sn_step_ex env_reward)."""
    from dataclasses import replace
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    T = 10
    base = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    specs = [replace(base, holding_cost=(0.5 + 0.25 * i,) * 4, backorder_cost=(1.0 + i,) * 4) for i in range(n_items)]
    rs = np.random.RandomState(4)
    init = rs.randint(0, 9, (n, n_items, 19))
    demand = rs.randint(0, 9, (n, n_items, T + 3))
    acts = rs.randint(0, 17, (T, n, n_items * 4))
    with _Kernel(which):
        mi = MultiItemSupplyNetwork(specs, n, "cuda", "int32")
        mi.load_episode(init, demand)
        mi.reset()
        r64 = torch.zeros((n * n_items,), dtype=torch.float64, device="cuda")
        for t in range(T):
            # reference value: per-chain double rewards of the same launch, summed in item order on the host
            a = torch.as_tensor(acts[t].astype(np.int32)).cuda()
            snap = mi.sim.get_state()
            mi.sim.step(a.view(n * n_items, 4), reward64_out=r64)
            per_chain = r64.cpu().numpy().reshape(n, n_items)
            want = per_chain[:, 0].copy()
            for i in range(1, n_items):
                want = want + per_chain[:, i]
            mi.sim.set_state(snap)
            _, rew, done = mi.step(a)
            assert np.array_equal(rew.cpu().numpy(), want.astype(np.float32)), t
        assert done


def test_multi_item_graph_step_equals_eager():
    from dataclasses import replace
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    n, T = 4096, 15
    base = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    specs = [base, replace(base, holding_cost=(0.75,) * 4)]
    rs = np.random.RandomState(8)
    init, demand = rs.randint(0, 9, (n, 2, 19)), rs.randint(0, 9, (n, 2, T + 3))
    acts = torch.as_tensor(rs.randint(0, 17, (T, n, 8)).astype(np.int32)).cuda()
    runs = []
    for use_graph in (False, True):
        mi = MultiItemSupplyNetwork(specs, n, "cuda", "int32", use_graph=use_graph)
        mi.load_episode(init, demand)
        mi.reset()
        tr = []
        for t in range(T):
            obs, rew, done = mi.step(acts[t])
            tr.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done))
        runs.append(tr)
        assert done
    for (o1, r1, d1), (o2, r2, d2) in zip(*runs):
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and d1 == d2


def test_int32_range_guard_flags_wrapped_quantities():
    """Orders beyond 2^26 are outside the exact int32 range: the kernels count the env-steps (stats[13]) and raise
    SN_FLAG_QTY_RANGE; ordinary runs leave both untouched."""
    from multi_echelon_drl_b200 import _lib
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    n, T = 500, 6
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 2, 0, 17)
    for which in ("plain", "tiled"):
        with _Kernel(which):
            sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
            sim.track_stats = True
            sim.load_episode(init, demand)
            sim.reset()
            sim.step(torch.as_tensor(acts[0]))
            assert sim.stats[13].item() == 0
            big = acts[1].copy()
            big[7, 2] = 1 << 30                                   # two such orders would wrap a backlog
            big[9, 0] = -(1 << 27)                                # |a| beyond the bound (clamped to 0 by the env)
            sim.step(torch.as_tensor(big))
            assert sim.stats[13].item() == 2
            sim.step(torch.as_tensor(acts[2]))                    # the big order is still in the pipeline / backlog
            assert sim.stats[13].item() >= 3
    # fused rollouts report envs, and the device-period launch raises the sticky flag
    sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
    sim.track_stats = True
    sim.load_episode(init, demand)
    a = acts.copy()
    a[1, 7, 2] = 1 << 30
    sim.episode(T, actions=torch.as_tensor(a))
    assert sim.stats[13].item() == 1
    sim.stats.zero_()
    sim.episode(T, actions=torch.as_tensor(acts))
    assert sim.stats[13].item() == 0
    sim.reset()
    d = torch.as_tensor(a[1].astype(np.int32)).cuda()
    sim.launch_step_device_period(d, sim.obs, sim.reward)
    assert sim.device_flags() & _lib.SN_FLAG_QTY_RANGE


@pytest.mark.parametrize("mode,n,dtype", [("onepass", 4096, "float32"), ("onepass", 128 * 148 - 5, "float32"),
                                          ("two_kernels", 4096, "float32"), ("two_kernels", 128 * 148 + 300, "float32"),
                                          ("two_kernels", 3000, "float64")])
def test_fused_vecnormalize_equals_eager_kernels(mode, n, dtype):
    """VecNormalize inside the replayed step graph gives what the three eager kernels give: identical observations /
    rewards from the env, statistics equal to 1e-12 relative (atomic accumulation orders differ), normalised outputs
    to float32 rounding; across two auto-resets.  onepass: statistics, grid barrier and normalisation in the step
    kernel itself (cooperative launch, one tile per SM); two_kernels: statistics fused, sn_vecnorm_apply behind it."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    envs = [VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, seed=4, dtype=dtype, use_graph=ug),
                         clip_obs=100.0, clip_reward=1000.0) for ug in (False, True)]
    envs[1].no_onepass = mode == "two_kernels"
    assert envs[1].graph_key()[2] == (mode == "onepass")
    o = [e.reset() for e in envs]
    assert torch.equal(envs[0].get_original_obs(), envs[1].get_original_obs())
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    graphed = 0
    for t in range(230):
        a = torch.rand((n, 4), device="cuda", generator=g) * 16
        res = [e.step(a.clone()) for e in envs]
        graphed += int(envs[1].venv.last_step_graphed)
        assert torch.equal(envs[0].get_original_obs(), envs[1].get_original_obs()), t
        assert torch.equal(envs[0].get_original_reward(), envs[1].get_original_reward()), t
        np.testing.assert_allclose(res[1][0].cpu().numpy(), res[0][0].cpu().numpy(), rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(res[1][1].cpu().numpy(), res[0][1].cpu().numpy(), rtol=2e-6, atol=1e-7)
        assert torch.equal(res[0][2], res[1][2])
        if res[0][2].all():
            np.testing.assert_allclose(res[1][3].terminal_observation.cpu().numpy(),
                                       res[0][3].terminal_observation.cpu().numpy(), rtol=2e-6, atol=2e-6)
    assert graphed == 230 - 2                                    # every step but the two terminal ones
    for k in ("mean", "var", "count"):
        np.testing.assert_allclose(envs[1].obs_rms[k].cpu().numpy(), envs[0].obs_rms[k].cpu().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(envs[1].ret_rms[k].cpu().numpy(), envs[0].ret_rms[k].cpu().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(envs[1].returns.cpu().numpy(), envs[0].returns.cpu().numpy(), rtol=1e-12, atol=1e-12)
    assert envs[1].venv.sim.device_flags() == 0
    # evaluation mode: frozen statistics, normalisation still inside the graph
    for e in envs:
        e.training = False
    before = envs[1]._obs_rms.clone()
    a = torch.rand((n, 4), device="cuda", generator=g) * 16
    r0, r1 = envs[0].step(a.clone()), envs[1].step(a.clone())
    assert torch.equal(before, envs[1]._obs_rms) and envs[1].venv.last_step_graphed
    np.testing.assert_allclose(r1[0].cpu().numpy(), r0[0].cpu().numpy(), rtol=2e-6, atol=2e-6)


def test_graph_vecenv_equals_eager_vecenv_without_vecnormalize():
    from multi_echelon_drl_b200 import register_envs as R
    n = 2000
    envs = [R.make("BeerGameNormalRetailer-v0", n_envs=n, seed=9, dtype="float32", use_graph=ug) for ug in (False, True)]
    o = [e.reset() for e in envs]
    assert torch.equal(o[0], o[1])
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for t in range(205):
        a = torch.rand((n, 1), device="cuda", generator=g) * 20
        (o0, r0, d0, i0), (o1, r1, d1, i1) = envs[0].step(a), envs[1].step(a)
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1), t
        if d0.all():
            assert torch.equal(i0.terminal_observation, i1.terminal_observation)
            assert torch.equal(i0.episode_return, i1.episode_return)


@pytest.mark.parametrize("n", [300, 1000, 128 * 148 + 37])
def test_tiled_writes_nothing_outside_its_buffers(n):
    """Canaries around every buffer the tiled kernel stores to (the TMA box of the last tile reaches beyond `ld`: the
    tensor map must clip it), padding columns n..ld-1 of the state included (compute-sanitizer is closed on this pool)."""
    from multi_echelon_drl_b200 import _lib
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    T = 6
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    init, demand, acts = _inputs(sp, n, T, 17, 0, 17)
    ld = (n + 31) // 32 * 32
    pad = 4096
    with _Kernel("tiled"):
        sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
        sim.load_episode(init, demand)

        def guarded(numel, dtype, fill):
            store = torch.full((numel + 2 * pad,), fill, dtype=dtype, device="cuda")
            return store, store[pad:pad + numel]

        st_store, st = guarded(40 * ld, torch.int32, -77)
        obs_store, obs = guarded(n * 28, torch.float32, -7.0)
        rew_store, rew = guarded(n, torch.float32, -7.0)
        ret_store, ret = guarded(n, torch.float64, -7.0)
        sim.reset()
        st.view(40, ld).copy_(sim.state)
        st.view(40, ld)[:, n:] = -55                                   # padding columns: must come back untouched
        ret.zero_()
        a = torch.as_tensor(acts[0].astype(np.int32)).cuda()
        io = _lib.SnStepIo()
        io.state, io.actions, io.demand, io.period = st.data_ptr(), a.data_ptr(), sim.demand[1].data_ptr(), 0
        io.obs, io.reward, io.ep_return = obs.data_ptr(), rew.data_ptr(), ret.data_ptr()
        _lib.check(sim.lib.sn_step_ex(C.byref(sim._c_spec), sim.dtype_code, n, ld, C.byref(io),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        want_obs, want_rew, _ = sim.step(a)                            # the same period on the simulator's own buffers
        assert torch.equal(obs.view(n, 28), want_obs) and torch.equal(rew, want_rew)
        assert torch.equal(st.view(40, ld)[:, :n], sim.state[:, :n]) and torch.equal(ret, sim.ep_return)
        assert (st.view(40, ld)[:, n:] == -55).all()
        for store, fill in ((st_store, -77), (obs_store, -7.0), (rew_store, -7.0), (ret_store, -7.0)):
            assert (store[:pad] == fill).all() and (store[-pad:] == fill).all()


def test_tiled_equals_plain_at_hbm_scale():
    """4 Mi + 77 envs (221 tiles per CTA, the size of the bench's `single_step_4M` line): state, observation and reward of
    three periods, tiled vs plain, bit for bit; inputs drawn on the device (host generation would dominate)."""
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    n, T = 4 * 1024 * 1024 + 77, 8
    sp = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    ld = (n + 31) // 32 * 32
    init = torch.randint(0, 9, (19, ld), generator=g, device="cuda", dtype=torch.int32)
    demand = torch.randint(0, 9, (T + 3, ld), generator=g, device="cuda", dtype=torch.int32)
    acts = torch.randint(-1, 18, (3, n, 4), generator=g, device="cuda", dtype=torch.int32)
    res = []
    for which in ("plain", "tiled"):
        with _Kernel(which):
            sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
            sim.track_stats = True
            sim.init.copy_(init); sim.demand.copy_(demand)
            sim.reset()
            sums = []
            for t in range(3):
                obs, rew, _ = sim.step(acts[t])
                sums.append((obs.clone(), rew.clone()))
            res.append((sums, sim.state.clone(), sim.stats.clone(), sim.ep_return.clone()))
            del sim
    for (o0, r0), (o1, r1) in zip(res[0][0], res[1][0]):
        assert torch.equal(o0, o1) and torch.equal(r0, r1)
    assert torch.equal(res[0][1][:, :n], res[1][1][:, :n]) and torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])


@pytest.mark.parametrize("kind,dtype", [("leadtimes", "int32"), ("leadtimes", "float32"), ("short_chain", "int32"),
                                        ("tree", "int32"), ("tree", "float64"), ("star", "int32")])
def test_tiled_equals_plain_on_other_networks(kind, dtype):
    """The tiled period step also serves chains with other lead times / fewer facilities (57 state words) and
    distribution trees of up to four nodes (64 words, one demand row per source): same results as the plain kernel."""
    from dataclasses import replace
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    T, n = 20, 128 * 148 + 4133
    base = _spec("uniform", (0, 1, 2, 3), True, "a", T)
    if kind == "leadtimes":
        sp = replace(base, info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1))
    elif kind == "short_chain":
        sp = replace(_spec("uniform", (0, 1), False, "a", T), n_facilities=3, info_leadtime=(2, 1, 2), ship_leadtime=(1, 2, 3),
                     holding_cost=(0.5,) * 3, backorder_cost=(1.0,) * 3, coplayer_target=(19, 20, 14))
    elif kind == "tree":
        sp = tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20), dict(name="dist_b", demand=(0, 5), target=14),
                        dict(name="maker", target=25)],
                       [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)],
                       ["dist_a", "dist_b"], True, max_episode_steps=T)
    else:
        sp = tree_spec([dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=12), dict(name="r3", demand=(0, 3), target=9),
                        dict(name="hub", target=40)],
                       [("ext", "hub", 2, 3), ("hub", "r3", 1, 1), ("hub", "r1", 2, 2), ("hub", "r2", 2, 1)], ["r1", "r2", "r3", "hub"], False,
                       max_episode_steps=T)
    init, demand = bulk_episode_inputs(sp, n, np.random.RandomState(31))
    acts = np.random.RandomState(32).randint(-1, 18, (T, n, sp.n_agents)).astype(np.float64)
    if dtype != "int32":
        acts = acts + np.random.RandomState(33).rand(*acts.shape)
    res = []
    for which in ("plain", "tiled"):
        os.environ["SN_NO_SPECIALIZED"] = "1"                       # the general library's kernels on both sides
        try:
            with _Kernel(which):
                sim = BatchedSupplyNetwork(sp, n, "cuda", dtype)
                sim.track_stats = True
                sim.load_episode(init, demand)
                sim.reset()
                r64 = torch.zeros((n,), dtype=torch.float64, device="cuda")
                cost = torch.zeros((8, sim.ld), dtype=torch.float64, device="cuda")
                tr = []
                for t in range(T):
                    obs, rew, done = sim.step(torch.as_tensor(acts[t]), reward64_out=r64, cost_out=cost)
                    tr.append((obs.clone(), r64.clone(), cost.clone()))
                res.append((tr, sim.state.clone(), sim.ep_return.clone(), sim.stats.clone(), done))
        finally:
            os.environ.pop("SN_NO_SPECIALIZED", None)
    for (o0, r0, c0), (o1, r1, c1) in zip(res[0][0], res[1][0]):
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(c0[:, :n], c1[:, :n])
    assert torch.equal(res[0][1][:, :n], res[1][1][:, :n]) and torch.equal(res[0][2], res[1][2]) and res[0][4] and res[1][4]
    if dtype == "int32":
        assert torch.equal(res[0][3], res[1][3])
    else:
        np.testing.assert_allclose(res[0][3].cpu().numpy(), res[1][3].cpu().numpy(), rtol=1e-12)
