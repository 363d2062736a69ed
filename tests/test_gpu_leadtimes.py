"""Generic lead-time kernels (serial chains with per-arc information / shipment lead times other than the
beer game's) vs golden trajectories recorded from the reference's own Node/Arc/SupplyNetwork classes and vs
the oracle on batches.  Integer quantities: bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_oracle_golden import CHAIN_CASES, LEADTIME_CASES, chain_spec, oracle_leadtime_rollout  # noqa: E402


def _sim(case, n, dtype="int32"):
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    return BatchedSupplyNetwork(chain_spec(case), n, "cuda", dtype)


def _init_for(sim, case, init36):
    """golden inits are stored in the 36-word layout; the beer game's own lead times use the packed 19."""
    if sim.init_words == 36:
        return init36
    cols = list(range(4))
    for k in range(4):
        cols += [4 + 4 * k + j for j in range(case["li"][k])]
    for k in range(4):
        cols += [20 + 4 * k + j for j in range(case["ls"][k])]
    return init36[:, cols]


@pytest.mark.parametrize("case", LEADTIME_CASES + CHAIN_CASES, ids=[c["id"] for c in LEADTIME_CASES + CHAIN_CASES])
def test_step_and_rollout_match_reference_golden(case):
    sim = _sim(case, 1)
    assert sim.state_words == (40 if case["li"] == (2, 2, 2, 1) and case["ls"] == (2, 2, 2, 2) else 57)
    assert sim.obs_dim == case["obs0"].shape[0]
    sim.load_episode(_init_for(sim, case, case["init"][None]), case["demand"][None])
    assert np.array_equal(sim.reset().cpu().numpy()[0], case["obs0"])
    r64 = torch.zeros(1, dtype=torch.float64, device="cuda")
    for t in range(100):
        obs, _, done = sim.step(torch.as_tensor(case["actions"][t][None]), reward64_out=r64)
        assert np.array_equal(obs.cpu().numpy()[0], case["obs"][t]), t
        assert r64.item() == case["rew"][t], t
    assert done
    out = sim.episode(100, actions=torch.as_tensor(case["actions"][:, None, :]), record_obs=True, record_reward=True)
    assert np.array_equal(out["traj_obs"].cpu().numpy()[:, 0], case["obs"])
    assert sim.ep_return.item() == case["rew"].sum()


@pytest.mark.parametrize("li,ls,agents,glob,rule", [
    ((1, 1, 1, 1), (1, 1, 1, 1), (0, 1, 2, 3), True, "a"),
    ((4, 1, 2, 3), (1, 4, 3, 2), (0, 1, 2, 3), True, "d+a"),
    ((3, 2, 1, 2), (2, 3, 4, 1), (2,), False, "a"),
    ((1, 4, 1, 2), (4, 1, 4, 3), (1, 2, 3), True, "a"),
    ((2, 2, 1), (2, 2, 2), (0, 1, 2), True, "a"),          # three facilities
    ((3, 1), (2, 4), (1,), True, "d+a"),                   # two
    ((2,), (3,), (0,), False, "a"),                        # a single facility buying from the external supplier
])
def test_generic_batch_vs_oracle_and_policy_rollout(li, ls, agents, glob, rule):
    """3,001 envs (ragged) x 100: step kernel, fused rollout with trajectory, in-kernel base-stock policy."""
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    case = dict(li=li, ls=ls, agents=agents, glob=glob, rule=rule)
    n = 3001
    sim = _sim(case, n)
    init, demand = bulk_episode_inputs(sim.spec, n, np.random.RandomState(5))
    assert init.shape == (n, 36)
    acts = np.random.RandomState(6).randint(-1, 19, (100, n, len(agents)))
    obs0, obs, rew = oracle_leadtime_rollout(case, init, demand, acts)
    sim.load_episode(init, demand)
    assert np.array_equal(sim.reset().cpu().numpy(), obs0)
    r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(100):
        o, _, _ = sim.step(torch.as_tensor(acts[t]), reward64_out=r64)
        assert np.array_equal(o.cpu().numpy(), obs[t]), t
        assert np.array_equal(r64.cpu().numpy(), rew[t]), t
    out = sim.episode(100, actions=torch.as_tensor(acts), record_obs=True)
    assert np.array_equal(out["traj_obs"].cpu().numpy(), obs)
    assert np.array_equal(sim.ep_return.cpu().numpy(), rew.sum(0))
    # in-kernel policy == heuristic kernel + step kernel
    from multi_echelon_drl_b200.simulator import heuristic_actions
    if len(agents) == len(li) and glob:
        pol = BasePolicyParams([19, 20, 20, 14][:len(li)], 0, 16, "a")
        sim.episode(100, policy=pol, write_state=False)
        fused = sim.ep_return.clone()
        o = sim.reset()
        for t in range(100):
            o, _, _ = sim.step(heuristic_actions(pol, o))
        assert torch.equal(fused, sim.ep_return)


def test_vecenv_with_other_lead_times_and_unsupported_ones():
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    spec = replace(uniform_spec(None, True), info_leadtime=(1, 2, 3, 1), ship_leadtime=(3, 3, 2, 4))
    env = SupplyNetVecEnv(spec, 512, seed=3, dtype="int32")
    obs = env.reset()
    s = env.sim.state[:, :512].to(torch.int64)
    for k in range(4):                                   # reference invariant: on_order == sum(pipeline)
        on_order = s[7 * k + 3:7 * k + 7].sum(0)
        pipe = s[28 + 3 * k:31 + 3 * k].sum(0) + s[40 + 4 * k:44 + 4 * k].sum(0) + (s[7 * (k + 1) + 1] if k < 3 else s[56])
        assert torch.equal(on_order, pipe)
    for _ in range(150):
        obs, rew, dones, infos = env.step(torch.full((512, 4), 4, device="cuda"))
    assert torch.isfinite(obs).all() and env.episodes_done == 512
    with pytest.raises(NotImplementedError):
        SupplyNetVecEnv(replace(uniform_spec(), info_leadtime=(3, 2, 2, 1), ship_leadtime=(3, 2, 2, 2)), 8)


def test_vecenv_on_a_two_facility_chain():
    """retailer + wholesaler in front of the external supplier: 14-number global observation, two actions,
    the same on-order invariant, and whole episodes through the in-kernel base-stock policy."""
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    u = uniform_spec(None, True)
    spec = replace(u, n_facilities=2, agents=(0, 1), info_leadtime=(2, 1), ship_leadtime=(2, 2), holding_cost=(0.5, 0.5),
                   backorder_cost=(1.0, 1.0), coplayer_target=(19.0, 20.0))
    env = SupplyNetVecEnv(spec, 700, seed=11, dtype="int32")
    obs = env.reset()
    assert obs.shape == (700, 14) and env.action_space.shape == (2,)
    s = env.sim.state[:, :700].to(torch.int64)
    for k in range(2):
        on_order = s[7 * k + 3:7 * k + 7].sum(0)
        pipe = s[28 + 3 * k:31 + 3 * k].sum(0) + s[40 + 4 * k:44 + 4 * k].sum(0) + s[7 * (k + 1) + 1]
        assert torch.equal(on_order, pipe)
    assert (s[14] >= (1 << 28)).all() and (s[21] >= (1 << 28)).all()      # phantom upstream stock
    assert not s[28 + 6:28 + 12].any() and not s[40 + 8:40 + 16].any()    # phantom pipelines stay empty
    total = torch.zeros(700, dtype=torch.float64, device="cuda")
    for _ in range(100):
        obs, rew, dones, infos = env.step(torch.full((700, 2), 4, device="cuda"))
        total += rew.double()
    assert dones.all() and (total < 0).all()
    stats = env.run_policy_episodes(BasePolicyParams([19, 20], 0, 16, "a"), n_episodes=2)
    stats = stats.cpu().numpy()
    assert stats[0] == 1400 * 100 and stats[10] == 1400 and stats[11] < 0 and stats[11] == stats[1]
    assert stats[4] == 0 and stats[5] == 0 and stats[8] == 0 and stats[9] == 0     # phantom nodes cost nothing
    with pytest.raises(ValueError):
        replace(spec, agents=(2,))


def test_multi_item_batches_on_other_lead_times_and_a_short_chain():
    """Items are independent copies of the chain with their own cost vectors (multi_item.py); with lead times other
    than the beer game's, and on a 3-facility chain, the interleaved batch equals one single-item run per item:
    same observations, reward = sum over items; stepwise and as one fused episode launch."""
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    u = uniform_spec(None, True)
    chains = [replace(u, info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1)),
              replace(u, n_facilities=3, agents=(0, 1, 2), info_leadtime=(2, 1, 2), ship_leadtime=(2, 3, 1), holding_cost=u.holding_cost[:3],
                      backorder_cost=u.backorder_cost[:3], coplayer_target=u.coplayer_target[:3])]
    for base in chains:
        nf, n = base.n_facilities, 1500
        items = [base, replace(base, holding_cost=(0.75, 0.3, 1.0, 0.25)[:nf], backorder_cost=(2.0, 1.0, 0.5, 3.0)[:nf]),
                 replace(base, holding_cost=(0.1,) * nf, backorder_cost=(4.0,) * nf)]
        mi = MultiItemSupplyNetwork(items, n)
        rs = np.random.RandomState(12)
        inputs = [bulk_episode_inputs(base, n, rs) for _ in items]
        init = np.stack([x[0] for x in inputs], axis=1)                  # [n, I, 36]
        dem = np.stack([x[1] for x in inputs], axis=1)                   # [n, I, 103]
        acts = rs.randint(-1, 19, (100, n, len(items) * nf))
        mi.load_episode(init, dem)
        singles = []
        for i, sp in enumerate(items):
            s = BatchedSupplyNetwork(sp, n)
            s.load_episode(*inputs[i])
            singles.append(s)
        o = mi.reset()
        for i, s in enumerate(singles):
            assert torch.equal(o[:, i * 7 * nf:(i + 1) * 7 * nf], s.reset())
        r64 = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in items]
        for t in range(100):
            o, r, done = mi.step(torch.as_tensor(acts[t]))
            tot = torch.zeros(n, dtype=torch.float64, device="cuda")
            for i, s in enumerate(singles):
                so, _, _ = s.step(torch.as_tensor(acts[t][:, i * nf:(i + 1) * nf]), reward64_out=r64[i])
                assert torch.equal(o[:, i * 7 * nf:(i + 1) * 7 * nf], so), (t, i)
                tot += r64[i]
            assert torch.equal(r, tot.float()), t
        assert done
        # the fused episode kernel picks the item's cost vector too
        a_all = torch.as_tensor(acts).reshape(100, n * len(items), nf)
        mi.sim.episode(100, actions=a_all)
        expect = torch.stack([s.ep_return for s in singles], dim=1).sum(1)
        assert torch.equal(mi.ep_return, expect)
