"""The reference's own integration tests (tests/test_supplynetwork.py:14-88), run against the B200
implementation through the single-env gym duck type the factories return when no batch size is given:

    env = make_beer_game(agent_managed_facilities=['retailer']); env.reset(); env.step(q)

Same order-quantity schedules, same 13 golden episode totals, and a gym-API conformance check standing
in for stable_baselines3's check_env (not installed): spaces, dtypes, shapes, 4-tuple step, terminal error.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ZERO, FOUR, TWELVE = [0] * 20, [4] * 20, [12] * 20
RAMP = [0, 0] + [8] * 18

# (agent-managed facilities, order quantities per step, expected episode total) — test_supplynetwork.py:25-69
GOLDEN = [
    (["retailer"], ZERO, -1656.0), (["retailer"], FOUR, -798.0), (["retailer"], TWELVE, -1260.0),
    (["manufacturer"], RAMP, -224), (["manufacturer"], FOUR, -644), (["manufacturer"], TWELVE, -658),
    (["distributor"], RAMP, -170), (["distributor"], FOUR, -784), (["distributor"], TWELVE, -644),
    (["wholesaler"], RAMP, -158.0), (["wholesaler"], FOUR, -864.0), (["wholesaler"], TWELVE, -888.0),
]


def play_twenty_periods(env, quantities):
    env.reset()
    total = 0
    for i in range(20):
        _, reward, _, _ = env.step(quantities[i])
        total += reward
    return total


@pytest.mark.parametrize("facilities,quantities,total", GOLDEN)
def test_classic_beer_game_single_facility_totals(facilities, quantities, total):
    from multi_echelon_drl_b200.envs import make_beer_game
    env = make_beer_game(agent_managed_facilities=facilities)
    q = np.array(quantities).reshape((20, 1))
    assert play_twenty_periods(env, q) == total
    assert play_twenty_periods(env, q) == total            # the same env object can be reset and replayed


def test_classic_beer_game_retailer_and_wholesaler():
    from multi_echelon_drl_b200.envs import make_beer_game
    env = make_beer_game(agent_managed_facilities=["retailer", "wholesaler"])
    q = np.array([0, 8] + [0, 4] + [0, 0] * 18).reshape(20, 2)
    assert play_twenty_periods(env, q) == -1656.0


def conforms_to_gym_api(env, n_agents, obs_dim):
    """What check_env verifies for these envs: spaces, reset/step signatures, dtypes, determinism of shapes."""
    assert env.observation_space.shape == (obs_dim,) and env.observation_space.dtype == np.float32
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float32 and obs.shape == (obs_dim,)
    for _ in range(3):
        a = env.action_space.sample()
        assert np.asarray(a).shape in ((n_agents,), ())
        out = env.step(np.atleast_1d(np.asarray(a)).astype(np.int64))
        assert isinstance(out, tuple) and len(out) == 4
        obs, reward, done, info = out
        assert obs.dtype == np.float32 and obs.shape == (obs_dim,) and np.all(np.isfinite(obs))
        assert isinstance(reward, float) and isinstance(done, bool) and isinstance(info, dict)
    assert env.seed(3) == [3]


@pytest.mark.parametrize("facilities", [["retailer"], ["wholesaler"], ["retailer", "wholesaler"],
                                        ["retailer", "wholesaler", "distributor", "manufacturer"]])
def test_gym_compatibility_of_the_three_factories(facilities):
    from multi_echelon_drl_b200.envs import (make_beer_game, make_beer_game_normal_multi_facility,
                                             make_beer_game_uniform_multi_facility)
    n = len(facilities)
    conforms_to_gym_api(make_beer_game(agent_managed_facilities=facilities), n, 7 * n)
    conforms_to_gym_api(make_beer_game_uniform_multi_facility(agent_managed_facilities=facilities, box_action_space=True), n, 7 * n)
    conforms_to_gym_api(make_beer_game_normal_multi_facility(agent_managed_facilities=facilities, box_action_space=True), n, 7 * n)


def test_single_env_errors_and_sn_view_match_reference_behaviour():
    from multi_echelon_drl_b200.envs import make_beer_game, make_beer_game_uniform_multi_facility
    from multi_echelon_drl_b200.utils.wrappers import wrap_action_d_plus_a  # noqa: F401  (vec env only)
    env = make_beer_game(agent_managed_facilities=["retailer"], max_period=3)
    env.reset()
    for _ in range(3):
        _, _, done, _ = env.step(np.array([4]))
    assert done and env.terminal and env.period == 3
    with pytest.raises(ValueError, match="terminal"):
        env.step(np.array([4]))                                              # envs.py:88-89
    env.reset()
    with pytest.raises(TypeError):
        env.step(["four"])                                                   # network_components.py:26-29
    with pytest.warns(RuntimeWarning, match="non-negative"):
        env.step(np.array([-3]))                                             # :368-374: warned and truncated to 0
    with pytest.raises(ValueError):
        make_beer_game_uniform_multi_facility(agent_managed_facilities=["retailer", "wholesaler"])   # needs box space
    # sn.get_state: the dict the d+a wrapper reads (utils/wrappers.py:24), classic initial state
    env = make_beer_game(agent_managed_facilities=["wholesaler"])
    env.reset()
    st = env.sn.get_state("wholesaler")
    assert list(st) == ["on_hand", "unfilled_demand", "latest_demand"] + [f"unreceived_pipeline_{i}" for i in range(4)]
    assert st == {"on_hand": 12, "unfilled_demand": 4, "latest_demand": 4, "unreceived_pipeline_0": 4,
                  "unreceived_pipeline_1": 4, "unreceived_pipeline_2": 4, "unreceived_pipeline_3": 4}
    assert env.sn.agent_managed_facilities == ["wholesaler"]
    # np.random.seed drives the single env like the reference (global generator), see seeded.npz
    import os
    from conftest import GOLDEN as GDIR
    z = np.load(os.path.join(GDIR, "seeded.npz"))
    env = make_beer_game_uniform_multi_facility(global_observable=True, box_action_space=True,
                                                agent_managed_facilities=["retailer", "wholesaler", "distributor", "manufacturer"])
    np.random.seed(12345)
    assert np.array_equal(env.reset(), z["uniform_12345_obs0"])


def test_registry_make_without_batch_returns_single_env_incl_dplusa():
    """gym.make(id) counterpart: single env; the DPlusA ids apply the d+a rule (register_envs.py:8-15)."""
    from multi_echelon_drl_b200 import register_envs as R
    from helpers import oracle_rollout
    env = R.make("BeerGameUniformMultiFacilityDPlusA-v0", seed=5)
    assert env.observation_space.shape == (28,) and env.max_episode_steps == 100 and env.spec.rule == "d+a"
    obs0 = env.reset()
    init, demand = env.sim.init[:, :1].t().cpu().numpy(), env.sim.demand[:, :1].t().cpu().numpy()
    acts = np.random.RandomState(0).randint(0, 17, (100, 1, 4))
    ref = oracle_rollout("uniform", (0, 1, 2, 3), False, "d+a", 100, init, demand, acts)
    assert np.array_equal(obs0, ref["obs0"][0])
    for t in range(100):
        obs, rew, done, _ = env.step(acts[t, 0])
        assert np.array_equal(obs, ref["obs"][t, 0]) and rew == ref["rew"][t, 0]
    assert done
    single = R.make("BeerGameNormalWholesalerFullInfo-v0")
    assert single.reset().shape == (28,) and single.action_space.shape == (1,)


def test_return_dict_observations_like_the_reference():
    """return_dict=True (envs.py:59,77-78): {agent: {on_hand, unfilled_demand, latest_demand, unreceived_pipeline_i}}."""
    from multi_echelon_drl_b200.envs import make_beer_game
    env = make_beer_game(agent_managed_facilities=["retailer", "wholesaler"], return_dict=True)
    obs = env.reset()
    assert list(obs) == ["retailer", "wholesaler"]
    assert obs["wholesaler"] == {"on_hand": 12, "unfilled_demand": 4, "latest_demand": 4, "unreceived_pipeline_0": 4,
                                 "unreceived_pipeline_1": 4, "unreceived_pipeline_2": 4, "unreceived_pipeline_3": 4}
    flat_env = make_beer_game(agent_managed_facilities=["retailer", "wholesaler"])
    flat = flat_env.reset()
    obs, r, d, _ = env.step(np.array([3, 5]))
    f2, r2, _, _ = flat_env.step(np.array([3, 5]))
    assert r == r2 and [v for a in obs.values() for v in a.values()] == f2.tolist()


@pytest.mark.parametrize("cid", [0, 1, 2])
def test_deterministic_random_demand_replays_the_reference_up_to_its_index_error(cid):
    """make_beer_game(demand_type='deterministic_random') (envs.py:519-520): one list of max_period values drawn from
    the global generator when the env is built (after np.random.seed(seed)), replayed by every episode.  The period
    advance reads one value more than the list holds, so the reference's episode dies at its last step with numpy's
    IndexError; golden episodes recorded from the live reference (tests/golden/make_deterministic_random.py)."""
    import os
    from multi_echelon_drl_b200.envs import make_beer_game
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "deterministic_random.npz"))
    p = f"c{cid}_"
    names = ["retailer", "wholesaler", "distributor", "manufacturer"]
    agents = [names[i] for i in z[p + "agents"]]
    T = int(z[p + "max_period"])
    env = make_beer_game(agents, demand_type="deterministic_random", max_period=T, seed=int(z[p + "seed"]))
    assert len(env.spec.demand_list) == T
    for episode in range(2):                                  # the same list again after a reset
        obs0 = env.reset()
        assert np.array_equal(obs0, z[p + "obs0"]) and np.array_equal(obs0, z[p + "obs0_again"])
        err_at = int(z[p + "err_at"])
        assert err_at == T - 1
        for t in range(err_at):
            o, r, d, info = env.step(z[p + "actions"][t])
            assert np.array_equal(o, z[p + "obs"][t]), (t, o, z[p + "obs"][t])
            assert r == z[p + "rew"][t] and not d
        with pytest.raises(IndexError, match=str(z[p + "err"])):
            env.step(z[p + "actions"][err_at])
    with pytest.raises(NotImplementedError):
        make_beer_game(agents, demand_type="deterministic_random", max_period=T, n_envs=8)
