"""Pins the CPU oracle (oracle/beergame_np.py) to the reference:
  * the 13 golden episode totals of the reference's own tests (tests/test_supplynetwork.py:25-69);
  * golden trajectories recorded from the unmodified reference (tests/golden/*.npz);
  * when /root/reference is mounted (build container), a live differential run.
Integer rollouts must be bit-exact (obs, reward, per-facility costs, done)."""
import numpy as np
import pytest

from conftest import CLASSIC_TOTALS, GOLDEN, load_golden_classic, load_golden_rollouts
from helpers import oracle_rollout
from oracle import beergame_np as O

GOLD = load_golden_rollouts()


def test_classic_totals_known_answers():
    eps = load_golden_classic()
    assert len(eps) == 13
    for ep, total in zip(eps, CLASSIC_TOTALS):
        ag = tuple(int(a) for a in ep["agents"])
        ref = oracle_rollout("classic", ag, False, "a", 35, ep["init"][None], ep["demand"][None], ep["actions"][:, None, :])
        assert np.array_equal(ref["obs0"][0], ep["obs0"])
        assert np.array_equal(ref["obs"][:, 0], ep["obs"])
        assert ref["rew"].sum() == total
        assert ep["rew"].sum() == total           # the fixture itself carries the reference's totals


@pytest.mark.parametrize("case", GOLD, ids=[c["id"] for c in GOLD])
def test_oracle_matches_golden_rollout(case):
    is_int = case["kind"] == "int"
    acts = case["actions"][:, None, :]
    ref = oracle_rollout(case["scenario"], case["agents"], case["global_observable"], case["rule"], 100,
                         case["init"][None], case["demand"][None], acts if is_int else acts.astype(np.float64),
                         dtype=np.int64 if is_int else np.float64)
    if is_int:
        assert np.array_equal(ref["obs0"][0], case["obs0"])
        assert np.array_equal(ref["obs"][:, 0], case["obs"])
        assert np.array_equal(ref["rew"][:, 0], case["rew"])
        assert np.array_equal(ref["holding"][:, 0], case["holding"])
        assert np.array_equal(ref["backorder"][:, 0], case["backorder"])
    else:
        # the reference runs these in mixed float32/python arithmetic (numpy >= 2 scalar rules);
        # 1e-6 relative plus an absolute term for cancellation residues of O(100) quantities
        np.testing.assert_allclose(ref["obs"][:, 0], case["obs"], rtol=1e-6, atol=2e-4)
        np.testing.assert_allclose(ref["rew"][:, 0], case["rew"], rtol=1e-6, atol=2e-3)
    assert ref["done"] and bool(case["done"][-1]) and not case["done"][:-1].any()


def test_oracle_heuristic_matches_golden():
    z = np.load(f"{GOLDEN}/heuristic.npz")
    for name, tg, ub in (("uniform", [19, 20, 20, 14], 16), ("normal", [48, 43, 41, 30], 20)):
        obs = z[f"{name}_obs"]
        for rule in ("a", "d+a"):
            got = O.base_stock_action(obs, tg, 0, ub, rule)
            np.testing.assert_allclose(got, z[f"{name}_{rule}_nd"], rtol=1e-6, atol=1e-5)
            np.testing.assert_allclose(got, z[f"{name}_{rule}_dict"], rtol=1e-6, atol=1e-5)
            assert np.array_equal(got[:32], z[f"{name}_{rule}_nd"][:32])       # integer rows exact
        got = O.base_stock_action(obs[:2], tg, 0, ub, "a", row0_broadcast=True)
        np.testing.assert_allclose(got[0], z[f"{name}_row0_batch2"].reshape(-1), rtol=1e-6)
        np.testing.assert_allclose(got[1], z[f"{name}_row0_batch2"].reshape(-1), rtol=1e-6)


def test_oracle_seeded_streams_match_golden():
    """np.random.seed(s); env.reset() draw order replayed by the oracle."""
    z = np.load(f"{GOLDEN}/seeded.npz")
    for scn in ("uniform", "normal"):
        for seed in (0, 1, 2, 12345, 2**31 - 1):
            inv0, so, sh, dem = O.seeded_episode_inputs(scn, seed)
            flat = list(inv0) + [x for k in range(4) for x in so[k]] + [x for k in range(4) for x in sh[k]]
            assert np.array_equal(np.array(flat), z[f"{scn}_{seed}_init"])
            assert np.array_equal(dem, z[f"{scn}_{seed}_demand"])


def test_oracle_batched_equals_single():
    """Vectorisation across envs does not couple them: a batch equals the single-env runs."""
    cases = [c for c in GOLD if c["kind"] == "int" and c["scenario"] == "uniform" and c["agents"] == (0, 1, 2, 3)
             and c["global_observable"] and c["rule"] == "a"]
    cases = cases + [dict(c, init=c["init"][::-1].copy()) for c in cases]     # a second, different env
    init = np.stack([c["init"] for c in cases])
    dem = np.stack([c["demand"] for c in cases])
    acts = np.stack([c["actions"] for c in cases], axis=1)
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", 100, init, dem, acts)
    assert np.array_equal(ref["obs"][:, 0], cases[0]["obs"])
    assert np.array_equal(ref["rew"][:, 0], cases[0]["rew"])
    assert not np.array_equal(ref["obs"][:, 1], cases[0]["obs"])


def test_oracle_rejects_bad_agent_sets_and_terminal_step():
    with pytest.raises(ValueError):
        O.BeerGameOracle(O.scenario_spec("uniform", (0, 2)), 1)
    orc = O.BeerGameOracle(O.scenario_spec("uniform", (0,), False, 2), 1)
    z = np.zeros((1, 2), int)
    orc.reset(np.zeros((1, 4), int), [z, z, z, z[:, :1]], [z] * 4, np.zeros((1, 5), int))
    orc.step(np.zeros((1, 1), int))
    _, _, done = orc.step(np.zeros((1, 1), int))
    assert done
    with pytest.raises(ValueError):
        orc.step(np.zeros((1, 1), int))


@pytest.mark.skipif(not __import__("oracle.ref_harness", fromlist=["x"]).reference_available(),
                    reason="/root/reference not mounted (GPU box)")
def test_oracle_vs_live_reference():
    from oracle import ref_harness as H
    rs = np.random.RandomState(99)
    for scn, hi in (("uniform", 17), ("normal", 21)):
        for ag in ((0, 1, 2, 3), (2,), (1, 2)):
            for glob in (False, True):
                seed = int(rs.randint(1 << 30))
                env = H.make_env(scn, [H.NODE_NAMES[k] for k in ag], glob, True, 100)
                acts = rs.randint(0, hi, (100, len(ag)))
                r = H.rollout(env, [a for a in acts], seed=seed, collect_costs=True)
                inv0, so, sh, dem = O.seeded_episode_inputs(scn, seed)
                flat = np.array(list(inv0) + [x for k in range(4) for x in so[k]] + [x for k in range(4) for x in sh[k]])
                ref = oracle_rollout(scn, ag, glob, "a", 100, flat[None], dem[None], acts[:, None, :])
                assert np.array_equal(ref["obs"][:, 0], r["obs"]) and np.array_equal(ref["rew"][:, 0], r["rew"])


def _leadtime_cases(fname="leadtimes.npz", prefix="g"):
    z = np.load(f"{GOLDEN}/{fname}")
    out = []
    for m in z["meta"]:
        cid, li, ls, ag, glob, rule = str(m).split("|")
        p = f"{prefix}{int(cid):03d}_"
        out.append(dict(id=f"li{li.replace(',', '')}-ls{ls.replace(',', '')}-ag{ag.replace(',', '')}-g{glob}-{rule}",
                        li=tuple(int(x) for x in li.split(",")), ls=tuple(int(x) for x in ls.split(",")),
                        agents=tuple(int(x) for x in ag.split(",")), glob=bool(int(glob)), rule=rule,
                        init=z[p + "init"], demand=z[p + "demand"], actions=z[p + "actions"], obs0=z[p + "obs0"],
                        obs=z[p + "obs"], rew=z[p + "rew"]))
    return out


LEADTIME_CASES = _leadtime_cases()
CHAIN_CASES = _leadtime_cases("chains.npz", "s")          # chains of 1..3 facilities


def oracle_leadtime_rollout(case, init36, demand, actions):
    """Oracle with per-arc lead times; init36 [n,36] = inv[4], so[k][0..3], sh[k][0..3]."""
    from dataclasses import replace
    nf = len(case["li"])
    spec = replace(O.scenario_spec("uniform", case["agents"], case["glob"], 100, case["rule"]), info_lt=case["li"],
                   ship_lt=case["ls"], n_facilities=nf)
    n = init36.shape[0]
    orc = O.BeerGameOracle(spec, n)
    so = [init36[:, 4 + 4 * k: 4 + 4 * k + case["li"][k]] for k in range(nf)]
    sh = [init36[:, 20 + 4 * k: 20 + 4 * k + case["ls"][k]] for k in range(nf)]
    obs0 = orc.reset(init36[:, :4], so, sh, demand)
    obs, rew = [], []
    for t in range(actions.shape[0]):
        o, r, _ = orc.step(actions[t])
        orc.check_invariant()
        obs.append(o); rew.append(r)
    return obs0, np.stack(obs), np.stack(rew)


@pytest.mark.parametrize("case", LEADTIME_CASES, ids=[c["id"] for c in LEADTIME_CASES])
def test_oracle_matches_golden_other_lead_times(case):
    """Serial chains with other per-arc lead times, recorded from the reference's own Node/Arc classes."""
    obs0, obs, rew = oracle_leadtime_rollout(case, case["init"][None], case["demand"][None], case["actions"][:, None, :])
    assert np.array_equal(obs0[0], case["obs0"])
    assert np.array_equal(obs[:, 0], case["obs"]) and np.array_equal(rew[:, 0], case["rew"])


@pytest.mark.parametrize("case", CHAIN_CASES, ids=[c["id"] for c in CHAIN_CASES])
def test_oracle_matches_golden_short_chains(case):
    """Chains of one to three facilities in front of the external supplier (reference Node/Arc classes)."""
    obs0, obs, rew = oracle_leadtime_rollout(case, case["init"][None], case["demand"][None], case["actions"][:, None, :])
    assert obs0.shape[1] == 7 * (len(case["li"]) if case["glob"] else len(case["agents"]))
    assert np.array_equal(obs0[0], case["obs0"])
    assert np.array_equal(obs[:, 0], case["obs"]) and np.array_equal(rew[:, 0], case["rew"])


def chain_spec(case, t_max=100):
    """Product ScenarioSpec of a lead-time / short-chain golden case."""
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    nf = len(case["li"])
    u = uniform_spec(None, case["glob"], t_max)
    spec = replace(u, n_facilities=nf, agents=case["agents"], info_leadtime=case["li"], ship_leadtime=case["ls"],
                   holding_cost=u.holding_cost[:nf], backorder_cost=u.backorder_cost[:nf],
                   coplayer_target=u.coplayer_target[:nf])
    return spec.with_rule("d+a", -8, 0, 16) if case["rule"] == "d+a" else spec


def test_seeded_streams_for_short_chains_match_reference():
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs, bulk_episode_inputs
    for cid, case in enumerate(CHAIN_CASES):
        spec = chain_spec(case)
        assert spec.obs_dim == case["obs0"].shape[0] and not spec.standard_leadtimes
        init, demand = seeded_episode_inputs(spec, [1300 + cid])
        assert np.array_equal(init[0], case["init"]), case["id"]
        assert np.array_equal(demand[0], case["demand"]), case["id"]
        bulk, _ = bulk_episode_inputs(spec, 64, np.random.RandomState(1))
        nf = len(case["li"])
        used = np.zeros(36, dtype=bool)
        used[:nf] = True
        for k in range(nf):
            used[4 + 4 * k: 4 + 4 * k + case["li"][k]] = True
            used[20 + 4 * k: 20 + 4 * k + case["ls"][k]] = True
        assert not bulk[:, ~used].any()


def test_seeded_streams_for_other_lead_times_match_reference():
    """The product's per-env stream generator replays the reference's RNG draw order for chains with other
    lead times too (number of shipment / sales-order draws per arc follows the arc's lead times)."""
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs
    z = np.load(f"{GOLDEN}/leadtimes.npz")
    for cid, case in enumerate(LEADTIME_CASES):
        spec = replace(uniform_spec(case["agents"], case["glob"], 100), info_leadtime=case["li"], ship_leadtime=case["ls"])
        init, demand = seeded_episode_inputs(spec, [900 + cid])
        if spec.standard_leadtimes:           # packed 19-word layout -> compare through the 36-word one
            full = np.zeros(36, dtype=np.int64)
            full[:4] = init[0, :4]
            off = 4
            for k in range(4):
                full[4 + 4 * k: 4 + 4 * k + case["li"][k]] = init[0, off: off + case["li"][k]]; off += case["li"][k]
            for k in range(4):
                full[20 + 4 * k: 22 + 4 * k] = init[0, off: off + 2]; off += 2
            got = full
        else:
            got = init[0]
        assert np.array_equal(got, case["init"]), case["id"]
        assert np.array_equal(demand[0], case["demand"]), case["id"]
