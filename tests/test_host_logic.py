"""Host-side logic that needs no GPU: scenario constants, seeded streams (RNG draw order of the
reference), Demand patterns (the reference's tests/test_demands.py, adapted), registry ids, the
lazy info list, the dict path of the heuristic."""
import numpy as np
import pytest

from conftest import GOLDEN


def test_scenario_constants_match_reference_factories():
    from multi_echelon_drl_b200 import classic_spec, normal_spec, uniform_spec
    u = uniform_spec()
    assert u.agents == (0, 1, 2, 3) and u.holding_cost == (0.5,) * 4 and u.backorder_cost == (1.0,) * 4
    assert u.coplayer_target == (19, 20, 20, 14) and (u.action_low, u.action_high) == (0, 16)       # envs.py:319-350,397-400
    assert u.init_inventory == (0, 25) and u.init_pipeline == (0, 9) and u.random_init              # envs.py:325-328
    nrm = normal_spec()
    assert nrm.agents == (0,) and nrm.holding_cost == (1.0, 0.75, 0.5, 0.25) and nrm.backorder_cost == (10.0, 0, 0, 0)
    assert nrm.coplayer_target == (48, 43, 41, 30) and nrm.action_high == 20 and not nrm.random_init  # envs.py:163-286
    assert nrm.coplayer_ub == 16                                                                      # Q6
    c = classic_spec()
    assert c.coplayer_target == (32, 32, 32, 24) and c.max_episode_steps == 35 and not c.random_init
    assert uniform_spec(["wholesaler"], False).obs_dim == 7 and uniform_spec(["wholesaler"], True).obs_dim == 28
    assert uniform_spec(["distributor", "wholesaler"]).agents == (2, 1)                                # caller order kept (Q14)
    with pytest.raises(ValueError):
        uniform_spec(["retailer", "distributor"])                                                      # Q13
    with pytest.raises(ValueError):
        uniform_spec(["retailer", "retailer"])
    with pytest.raises(ValueError):
        uniform_spec(["nobody"])


def test_seeded_streams_replay_reference_draw_order():
    """np.random.seed(s); env.reset() in the reference == seeded_episode_inputs(spec, [s])."""
    from multi_echelon_drl_b200 import normal_spec, uniform_spec
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs
    z = np.load(f"{GOLDEN}/seeded.npz")
    seeds = (0, 1, 2, 12345, 2**31 - 1)
    for name, spec in (("uniform", uniform_spec()), ("normal", normal_spec(random_init=True))):
        init, demand = seeded_episode_inputs(spec, seeds)
        for i, s in enumerate(seeds):
            assert np.array_equal(init[i], z[f"{name}_{s}_init"]), (name, s)
            assert np.array_equal(demand[i], z[f"{name}_{s}_demand"]), (name, s)
    # fixed-init variants draw only the demand
    init, demand = seeded_episode_inputs(uniform_spec(random_init=False), [5])
    assert init[0].tolist() == [12] * 4 + [4] * 15
    assert np.array_equal(demand[0], np.random.RandomState(5).randint(0, 9, 103))
    init, _ = seeded_episode_inputs(normal_spec(), [5])
    assert init[0].tolist() == [10] * 4 + [10, 0, 10, 0, 10, 0, 10] + [10, 0] * 4


def test_persistent_streams_continue_across_resets():
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.utils.demands import seeded_episode_inputs_from_streams
    spec = uniform_spec()
    streams = [np.random.RandomState(3)]
    a = seeded_episode_inputs_from_streams(spec, streams)
    b = seeded_episode_inputs_from_streams(spec, streams)
    rs = np.random.RandomState(3)
    n_draws = 122                                  # 4 inv + 103 demand + 8 ship + 7 SO
    first = [rs.randint(0, 25)] + list(rs.randint(0, 9, 103)) + [rs.randint(0, 25) for _ in range(3)] + \
            [rs.randint(0, 9) for _ in range(15)]
    assert len(first) == n_draws and first[0] == a[0][0, 0] and list(a[1][0]) == first[1:104]
    assert rs.randint(0, 25) == b[0][0, 0]         # second reset continues the same stream


def test_demand_patterns_like_reference_tests():
    """tests/test_demands.py of the reference: uniform hits both bounds inclusively; normal(10,4)
    moments and clipping at 0; classic pattern; list passthrough; argument validation."""
    from multi_echelon_drl_b200.utils.demands import Demand
    rs = np.random.RandomState(0)
    for lo, hi in ((0, 2), (0, 8)):
        d = Demand("uniform", low=lo, high=hi, size=100).sample(4, rs)
        assert d.shape == (4, 103) and d.min() == lo and d.max() == hi
    d = Demand("normal", mean=10, sd=4, size=100000).sample(1, rs)[0]
    assert abs(d.mean() - 10) < 0.1 and abs(d.std() - 4) < 0.04 and d.min() == 0
    assert Demand("classic_beer_game", size=35).sample(2, rs)[1].tolist() == [4] * 4 + [8] * 34
    assert Demand([1, 2, 3]).sample(2, rs).tolist() == [[1, 2, 3]] * 2
    with pytest.raises(ValueError):
        Demand("uniform")
    with pytest.raises(ValueError):
        Demand("normal", mean=1)
    with pytest.raises(ValueError):
        Demand("poisson")
    # one stream, same draws as the reference's np.random calls
    a = Demand("uniform", low=0, high=8, size=100).draw(np.random.RandomState(1))
    assert np.array_equal(a, np.random.RandomState(1).randint(0, 9, 103))
    b = Demand("normal", mean=10, sd=2, size=100).draw(np.random.RandomState(1))
    assert np.array_equal(b, np.round(np.maximum(10 + 2 * np.random.RandomState(1).randn(103), 0)).astype(int))


def test_registry_has_the_38_reference_ids():
    from multi_echelon_drl_b200 import register_envs as R
    ids = R.register_envs()
    assert len(ids) == 38
    for kind in ("Uniform", "Normal"):
        assert f"BeerGame{kind}MultiFacilityFullInfo-v0" in ids and f"BeerGame{kind}MultiFacilityDPlusA-v0" in ids
        for role in ("Retailer", "Wholesaler", "Distributor", "Manufacturer"):
            for suffix in ("", "Discrete", "FullInfo", "DiscreteFullInfo"):
                assert f"BeerGame{kind}{role}{suffix}-v0" in ids
    assert "BeerGameUniformRetailerDPlusA-v0" in ids and "BeerGameUniformRetailerDiscreteDPlusA-v0" in ids
    with pytest.raises(KeyError):
        R.make("BeerGameNope-v0")


def test_lazy_infos_behaves_like_a_list_of_dicts():
    from multi_echelon_drl_b200.vec_env import LazyInfos
    empty = LazyInfos(3)
    assert len(empty) == 3 and empty[0] == {} and empty[-1] == {} and list(empty) == [{}, {}, {}]
    with pytest.raises(IndexError):
        empty[3]
    obs = np.arange(6.0).reshape(3, 2)
    done = LazyInfos(3, obs, np.array([-1.0, -2.5, -3.0]), 100, 1.25)
    assert done[1]["episode"] == {"r": -2.5, "l": 100, "t": 1.25}
    assert np.array_equal(done[2]["terminal_observation"], obs[2]) and done[0]["TimeLimit.truncated"] is False
    assert len(done[0:2]) == 2


def test_shard_ranges_partition_the_batch():
    from multi_echelon_drl_b200.distributed import shard_range
    for n, w in ((1048576, 8), (65536, 4), (10, 3), (7, 8), (1, 1)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_parse_order_sequence_reference_cases():
    """The three cases of the reference's tests/test_graph.py:11-49."""
    from multi_echelon_drl_b200.utils.graph import parse_order_sequence
    assert parse_order_sequence(
        ["manufacturer", "distributor", "wholesaler", "retailer"],
        {"is_external_supplier": ["manufacturer"], "manufacturer": ["distributor"], "distributor": ["wholesaler"],
         "wholesaler": ["retailer"], "retailer": []}) == ["retailer", "wholesaler", "distributor", "manufacturer"]
    assert parse_order_sequence(
        ["1", "2", "3", "4", "5", "6", "7"],
        {"6": ["4"], "5": ["2", "7"], "4": ["3", "1"], "2": ["1"], "3": [], "1": ["7"], "7": []}) == \
        ["7", "1", "2", "3", "4", "5", "6"]
    assert parse_order_sequence(
        ["1", "2", "3", "4", "5", "6", "7", "8", "9"],
        {"9": ["3"], "7": ["6", "2"], "3": ["5", "1"], "6": ["1"], "5": [], "1": ["2"], "2": ["4"], "4": ["8"],
         "8": []}) == ["8", "4", "2", "1", "5", "3", "6", "7", "9"]


def test_yaml_configs_lower_to_specs():
    """tests/test_graph.py:52-63 (5 nodes, 4 arcs, info lead times 2 and 1) + a working loader."""
    import os
    from multi_echelon_drl_b200.spec import specs_from_config
    from multi_echelon_drl_b200.utils.graph import chain_from_config, read_yaml
    import multi_echelon_drl_b200 as pkg
    d = os.path.join(os.path.dirname(pkg.__file__), "network_configs")
    net = read_yaml(os.path.join(d, "beergame.yaml"))
    assert len(net["nodes"]) == 5 and len(net["arcs"]) == 4
    assert net["arcs"][0]["info_leadtime"] == 2 and net["arcs"][3]["info_leadtime"] == 1
    seq, info, ship, hold, back = chain_from_config(net)
    assert seq == ["retailer", "wholesaler", "distributor", "manufacturer"] and info == (2, 2, 2, 1) and ship == (2,) * 4
    (spec,) = specs_from_config(net, ["retailer", "wholesaler", "distributor", "manufacturer"], True)
    assert spec.holding_cost == (0.5,) * 4 and spec.backorder_cost == (1.0,) * 4 and spec.obs_dim == 28
    multi = specs_from_config(read_yaml(os.path.join(d, "multi_item.yaml")))
    assert len(multi) == 2 and multi[1].holding_cost == (0.75,) * 4 and multi[0].holding_cost == (0.5,) * 4
    other = read_yaml(os.path.join(d, "beergame.yaml"))
    other["arcs"][0]["info_leadtime"] = 3                # 3 + 2 <= 5: generic lead-time kernels
    (sp3,) = specs_from_config(other)
    assert sp3.info_leadtime == (3, 2, 2, 1) and not sp3.standard_leadtimes and spec.standard_leadtimes
    bad = read_yaml(os.path.join(d, "beergame.yaml"))
    bad["arcs"][0]["info_leadtime"] = 4                  # 4 + 2 > 5: the reference's own assertion fails there
    with pytest.raises(NotImplementedError):
        specs_from_config(bad)
    bad = read_yaml(os.path.join(d, "beergame.yaml"))
    bad["arcs"].append({"customer": "retailer", "supplier": "distributor", "info_leadtime": 1, "shipment_leadtime": 1})
    with pytest.raises(NotImplementedError):
        specs_from_config(bad)


def test_hge_rate_schedule_and_single_coin():
    torch = pytest.importorskip("torch")
    from multi_echelon_drl_b200.hge import HgeRateSchedule, hge_predict
    sch = HgeRateSchedule(0.5, 1000)
    assert sch(0) == 0.5 and abs(sch(50_000) - 0.25) < 1e-12 and sch(100_000) == 0.0 and sch(200_000) == 0.0

    class H:
        def get_order_quantity(self, o):
            return torch.full((o.shape[0], 4), 7.0)

    obs = torch.zeros(6, 28)
    pol = lambda o: torch.zeros(o.shape[0], 4)
    rng = np.random.RandomState(0)
    heads = sum(bool(hge_predict(pol, H(), obs, obs, 0.5, rng=rng)[1]) for _ in range(4000))
    assert abs(heads / 4000 - 0.25) < 0.03                       # probability hge_rate**2 per call (hge.py:157)
    a, used = hge_predict(pol, H(), obs, obs, 1.0, rng=rng)
    assert used is True and (a == 7).all()
    a, used = hge_predict(pol, H(), obs, obs, 1.0, deterministic=True)
    assert used is False and (a == 0).all()
    a, mask = hge_predict(pol, H(), obs, obs, 1.0, per_env=True)
    assert mask.all() and (a == 7).all()


def test_samples_demand_pattern_cycles_rows(tmp_path):
    """Demand('samples', data_path=...) (utils/demands.py:26-29,54-56): rows of the file, cyclically."""
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.utils.demands import Demand, bulk_episode_inputs, seeded_episode_inputs
    from dataclasses import replace
    rows = np.arange(3 * 103).reshape(3, 103) % 9
    path = str(tmp_path / "demand.npy")
    np.save(path, rows)
    d = Demand("samples", data_path=path, size=100)
    got = [d.draw(None) for _ in range(4)]
    assert np.array_equal(got[0], rows[0]) and np.array_equal(got[2], rows[2]) and np.array_equal(got[3], rows[0])
    spec = replace(uniform_spec(), demand_pattern="samples", demand_path=path)
    _, dem = bulk_episode_inputs(spec, 4, np.random.RandomState(0))
    assert np.array_equal(dem[0], rows[0]) and np.array_equal(dem[3], rows[0]) and dem.shape == (4, 103)
    _, dem2 = seeded_episode_inputs(spec, [1])
    assert np.array_equal(dem2[0], rows[1])                  # the file pointer keeps advancing across resets
    from multi_echelon_drl_b200.envs import make_beer_game_uniform_multi_facility
    with pytest.raises(FileNotFoundError):
        make_beer_game_uniform_multi_facility(env_mode="test", box_action_space=True)
    with pytest.raises(ValueError):
        make_beer_game_uniform_multi_facility(env_mode="validate", box_action_space=True)


def test_multi_item_spec_combination_rules():
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.multi_item import combine_item_specs
    a = uniform_spec(None, True)
    b = replace(a, holding_cost=(0.75,) * 4, backorder_cost=(2.0,) * 4)
    c = combine_item_specs([a, b])
    assert c.n_items == 2 and c.extra_item_costs == (((0.75,) * 4, (2.0,) * 4),)
    cs = c.to_c()
    assert cs.n_items == 2 and list(cs.item_holding_cost[0])[:4] == [0.75] * 4 and list(cs.item_backorder_cost[0])[:4] == [2.0] * 4
    assert list(cs.holding_cost)[:4] == [0.5] * 4 and list(cs.holding_cost)[4:] == [0.0] * 4       # arrays are 8 wide (ABI v7)
    with pytest.raises(ValueError):
        combine_item_specs([a, replace(a, coplayer_target=(1, 2, 3, 4))])       # only costs may differ
    with pytest.raises(ValueError):
        combine_item_specs([a] * 5)
    with pytest.raises(ValueError):
        combine_item_specs([])


def test_spec_lowering_to_c_struct_roundtrip():
    from dataclasses import replace
    from multi_echelon_drl_b200 import normal_spec
    s = replace(normal_spec(["wholesaler", "retailer"], True, 50, True), info_leadtime=(1, 2, 3, 1), ship_leadtime=(4, 3, 2, 2))
    s = s.with_rule("d+a", -10, 0, 20)
    c = s.to_c()
    assert (c.n_facilities, c.n_agents, list(c.agents)[:2], c.global_observable, c.max_episode_steps, c.rule) == (4, 2, [1, 0], 1, 50, 1)
    assert list(c.info_leadtime)[:4] == [1, 2, 3, 1] and list(c.ship_leadtime)[:4] == [4, 3, 2, 2]
    assert list(c.holding_cost)[:4] == [1.0, 0.75, 0.5, 0.25] and list(c.backorder_cost)[:4] == [10.0, 0, 0, 0]
    assert list(c.coplayer_target)[:4] == [48, 43, 41, 30] and (c.coplayer_lb, c.coplayer_ub) == (0.0, 16.0)
    assert len(c.info_leadtime) == 8 and not any(list(c.info_leadtime)[4:])
    assert (c.dpa_offset, c.dpa_lb, c.dpa_ub) == (-10.0, 0.0, 20.0) and c.n_items == 1
    assert s.obs_dim == 28 and s.obs_nodes == (0, 1, 2, 3) and not s.standard_leadtimes
    with pytest.raises(ValueError):
        replace(s, info_leadtime=(1, 2, 3))
    with pytest.raises(ValueError):
        s.with_rule("a+d", 0, 0, 1)


def test_init_layouts_for_standard_and_generic_lead_times():
    from dataclasses import replace
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs, init_layout, seeded_episode_inputs
    std = uniform_spec()
    gen = replace(std, info_leadtime=(1, 3, 2, 1), ship_leadtime=(4, 2, 3, 1))
    assert init_layout(std) == (19, (4, 6, 8, 10), (11, 13, 15, 17))
    assert init_layout(gen) == (36, (4, 8, 12, 16), (20, 24, 28, 32))
    init, _ = bulk_episode_inputs(gen, 50, np.random.RandomState(0))
    assert init.shape == (50, 36)
    for k, (li, ls) in enumerate(zip(gen.info_leadtime, gen.ship_leadtime)):     # slots beyond the lead time stay empty
        assert (init[:, 4 + 4 * k + li: 8 + 4 * k] == 0).all() and (init[:, 20 + 4 * k + ls: 24 + 4 * k] == 0).all()
    # per-env streams follow the reference's draw order with the arc's own number of draws
    init, demand = seeded_episode_inputs(gen, [7])
    rs = np.random.RandomState(7)
    inv_r = rs.randint(0, 25); dem = rs.randint(0, 9, 103); rest = [rs.randint(0, 25) for _ in range(3)]
    assert init[0, 0] == inv_r and np.array_equal(demand[0], dem) and init[0, 1:4].tolist() == rest
    for k in (3, 2, 1, 0):
        sh = [rs.randint(0, 9) for _ in range(gen.ship_leadtime[k])]
        so = [rs.randint(0, 9) for _ in range(gen.info_leadtime[k])]
        assert init[0, 20 + 4 * k: 20 + 4 * k + len(sh)].tolist() == sh and init[0, 4 + 4 * k: 4 + 4 * k + len(so)].tolist() == so


def test_short_chain_specs_and_yaml():
    """Chains of 1..3 facilities: spec validation, observation layout, YAML factory, C-side validation."""
    import ctypes as C
    from dataclasses import replace
    from multi_echelon_drl_b200 import _lib, uniform_spec
    from multi_echelon_drl_b200.spec import specs_from_config
    u = uniform_spec(None, True)
    s2 = replace(u, n_facilities=2, agents=(1, 0), info_leadtime=(2, 1), ship_leadtime=(2, 2))
    assert s2.obs_nodes == (0, 1) and s2.obs_dim == 14 and s2.node_names == ("retailer", "wholesaler")
    assert not s2.standard_leadtimes
    assert replace(s2, global_observable=False).obs_nodes == (1, 0)
    with pytest.raises(ValueError):
        replace(u, n_facilities=2, agents=(0, 2))
    with pytest.raises(ValueError):
        replace(u, n_facilities=5)
    with pytest.raises(ValueError):
        replace(u, n_facilities=3, agents=(0,), info_leadtime=(2, 2))
    lib = _lib.load()
    c = s2.to_c()
    assert c.n_facilities == 2 and lib.sn_spec_validate(C.byref(c), _lib.SN_I32) == 0
    assert lib.sn_obs_dim(C.byref(c)) == 14 and lib.sn_state_words(C.byref(c)) == 57 and lib.sn_init_words(C.byref(c)) == 36
    c.agents[0] = 2
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) != 0 and b"agents" in lib.sn_last_error()
    cfg = dict(nodes=[dict(name="retailer", holding_cost=0.5, backorder_cost=1, demand_path="uniform:0:8"),
                      dict(name="wholesaler", holding_cost=0.25, backorder_cost=2),
                      dict(name="external_supplier", is_external_supplier=True)],
               arcs=[dict(customer="wholesaler", supplier="external_supplier", info_leadtime=1, shipment_leadtime=3),
                     dict(customer="retailer", supplier="wholesaler", info_leadtime=2, shipment_leadtime=2)])
    (sp,) = specs_from_config(cfg, ["wholesaler"], True, 100)
    assert sp.n_facilities == 2 and sp.agents == (1,) and sp.info_leadtime == (2, 1) and sp.ship_leadtime == (2, 3)
    assert sp.holding_cost == (0.5, 0.25) and sp.backorder_cost == (1.0, 2.0) and sp.obs_dim == 14
    assert sp.coplayer_target == (19, 20)
    with pytest.raises(ValueError):               # no demand source left (other names go down the tree path, which says so)
        specs_from_config(dict(nodes=cfg["nodes"][1:], arcs=cfg["arcs"][:1]), ["wholesaler"])


def test_bind_to_gpu_cpus_is_a_no_op_without_a_gpu():
    """bench.py pins each rank to its GPU's CPUs before allocating pinned buffers; without NVML / a GPU (this
    container) or on single-node VMs nothing changes and None is returned."""
    import os
    from multi_echelon_drl_b200.distributed import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_cpus(0) is None or isinstance(bind_to_gpu_cpus(0), list)
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_list_pattern_rows_are_padded_to_the_episode_size_and_replayed_every_episode():
    """utils/demands.py:38-39: a list / array pattern is the episode's demand as is.  make_beer_game's
    'deterministic_random' builds one of max_period values (envs.py:519-520), three short of what Demand.size says;
    the batched table row is zero-padded and the single env raises the reference's IndexError before the padding is
    read (tests/test_gpu_reference_suite.py)."""
    from multi_echelon_drl_b200 import spec as S
    from multi_echelon_drl_b200.utils import demands as D
    np.random.seed(3)
    lst = (8 + 2 * np.random.randn(35)).astype(int)
    sp = S.classic_spec(["retailer"], 35, "list", lst)
    assert sp.demand_pattern == "list" and sp.demand_list == tuple(int(x) for x in lst) and hash(sp) is not None
    dem = D.demand_for(sp)
    assert dem.size == 38
    state = np.random.get_state()[1].copy()
    for _ in range(2):
        init, rows = D.seeded_episode_inputs_from_streams(sp, [np.random])
        assert rows.shape == (1, 38) and np.array_equal(rows[0, :35], lst) and not rows[0, 35:].any()
    assert np.array_equal(np.random.get_state()[1], state)          # a list pattern draws nothing at reset
    assert np.array_equal(dem.sample(3, np.random), np.broadcast_to(lst, (3, 35)))      # passed through as is
    long = D.Demand(list(range(50)), size=10)
    assert np.array_equal(long.draw(np.random), np.arange(50)) and np.array_equal(long.row(np.random), np.arange(13))
