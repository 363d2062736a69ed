"""Distribution-tree oracle (oracle/network_np.py) vs trajectories recorded from the reference's own classes on
tree networks (tests/golden/trees.npz), vs the serial-chain oracle, and - when /root/reference is mounted -
vs the live reference."""
import json
import os

import numpy as np
import pytest

from oracle.network_np import EXTERNAL, TreeOracle, TreeSpec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases(fname="trees.npz", prefix="t"):
    z = np.load(f"{GOLDEN}/{fname}")
    out = []
    for m in z["meta"]:
        c = json.loads(str(m))
        p = f"{prefix}{c['cid']:03d}_"
        c["id"] = f"{c['topo']}-ag{''.join(map(str, c['agents']))}-g{int(c['glob'])}-{c['rule']}"
        for k in ("init", "actions", "obs0", "obs", "rew"):
            c[k] = z[p + k]
        c["demand_table"] = z[p + "demand"]                       # [n_sources, T+3]; c["demand"] stays the per-node ranges
        c["ref_holding"], c["ref_backorder"] = z[p + "holding"], z[p + "backorder"]   # c["holding"] = coefficients
        out.append(c)
    return out


TREE_CASES = _cases()
BIG_TREE_CASES = _cases("trees8.npz", "b")             # five to eight internal nodes, 72-word initial-state layout


def init_layout(n_nodes):
    """(inventory rows, first sales-order row, first shipment row) of the init table: 36 words up to four nodes, 72 beyond."""
    return (8, 8, 40) if n_nodes > 4 else (4, 4, 20)


def oracle_spec(case) -> TreeSpec:
    return TreeSpec(names=case["names"], supplier=case["supplier"], demand_source=[d is not None for d in case["demand"]],
                    info_lt=case["li"], ship_lt=case["ls"], holding_cost=case["holding"], backorder_cost=case["backorder"],
                    coplayer_target=case["target"], arc_order=case["arc_order"], agents=tuple(case["agents"]),
                    global_observable=case["glob"], rule=case["rule"])


def tree_oracle_rollout(case, init36, demand, actions, return_costs=False):
    """init36 [n, 36 | 72] (inv, so[k][0..3], sh[k][0..3] by node); demand [n, n_sources, T+3]; actions [T, n, n_agents]."""
    spec = oracle_spec(case)
    n, N = init36.shape[0], len(case["names"])
    orc = TreeOracle(spec, n)
    _, so0, sh0 = init_layout(N)
    so = [init36[:, so0 + 4 * k: so0 + 4 * k + case["li"][k]] for k in range(N)]
    sh = [init36[:, sh0 + 4 * k: sh0 + 4 * k + case["ls"][k]] for k in range(N)]
    obs0 = orc.reset(init36[:, :N], so, sh, demand)
    obs, rew, hold, back = [], [], [], []
    for t in range(actions.shape[0]):
        o, r, _ = orc.step(actions[t])
        orc.check_invariant()
        obs.append(o); rew.append(r); hold.append(orc.holding.copy()); back.append(orc.backorder.copy())
    if return_costs:
        return obs0, np.stack(obs), np.stack(rew), np.stack(hold), np.stack(back)
    return obs0, np.stack(obs), np.stack(rew)


@pytest.mark.parametrize("case", TREE_CASES + BIG_TREE_CASES, ids=[c["id"] for c in TREE_CASES + BIG_TREE_CASES])
def test_tree_oracle_matches_reference_golden(case):
    obs0, obs, rew, hold, back = tree_oracle_rollout(case, case["init"][None], case["demand_table"][None],
                                                     case["actions"][:, None, :], return_costs=True)
    spec = oracle_spec(case)
    assert [case["names"][k] for k in spec.sequence] == case["order_sequence"]
    assert np.array_equal(obs0[0], case["obs0"])
    assert np.array_equal(obs[:, 0], case["obs"]) and np.array_equal(rew[:, 0], case["rew"])
    assert np.array_equal(hold[:, 0], case["ref_holding"]) and np.array_equal(back[:, 0], case["ref_backorder"])


def test_tree_oracle_equals_chain_oracle_on_chains():
    """A serial chain is a tree: both restatements agree on a batch, for two lead-time sets and agent sets."""
    from dataclasses import replace
    from oracle import beergame_np as O
    rs = np.random.RandomState(3)
    for li, ls, ag, glob, rule in (((2, 2, 2, 1), (2, 2, 2, 2), (0, 1, 2, 3), True, "a"), ((3, 1, 2, 2), (2, 4, 3, 1), (2, 1), False, "d+a"),
                                   ((1, 2, 2), (2, 2, 3), (1,), True, "a")):
        nf, n = len(li), 64
        init = np.zeros((n, 36), dtype=np.int64)
        init[:, :nf] = rs.randint(0, 25, (n, nf))
        for k in range(nf):
            init[:, 4 + 4 * k: 4 + 4 * k + li[k]] = rs.randint(0, 9, (n, li[k]))
            init[:, 20 + 4 * k: 20 + 4 * k + ls[k]] = rs.randint(0, 9, (n, ls[k]))
        dem = rs.randint(0, 9, (n, 103))
        acts = rs.randint(-1, 19, (100, n, len(ag)))
        cs = replace(O.scenario_spec("uniform", ag, glob, 100, rule), info_lt=li, ship_lt=ls, n_facilities=nf)
        chain = O.BeerGameOracle(cs, n)
        o0 = chain.reset(init[:, :4], [init[:, 4 + 4 * k: 4 + 4 * k + li[k]] for k in range(nf)],
                         [init[:, 20 + 4 * k: 20 + 4 * k + ls[k]] for k in range(nf)], dem)
        names = ["retailer", "wholesaler", "distributor", "manufacturer"][:nf]
        case = dict(names=names, supplier=list(range(1, nf)) + [EXTERNAL], demand=[(0, 8)] + [None] * (nf - 1), li=li, ls=ls,
                    holding=[0.5] * nf, backorder=[1.0] * nf, target=[19, 20, 20, 14][:nf], arc_order=list(range(nf - 1, -1, -1)),
                    agents=ag, glob=glob, rule=rule)
        t0, tobs, trew = tree_oracle_rollout(case, init, dem[:, None, :], acts)
        assert np.array_equal(o0, t0)
        for t in range(100):
            o, r, _ = chain.step(acts[t])
            assert np.array_equal(o, tobs[t]) and np.array_equal(r, trew[t]), t


def test_tree_spec_rejects_agents_not_contiguous_in_order_sequence():
    c = TREE_CASES[0]
    with pytest.raises(ValueError):
        TreeSpec(names=["shop", "dist_a", "dist_b", "maker"], supplier=[1, 3, 3, EXTERNAL], demand_source=[True, False, True, False],
                 info_lt=[2, 2, 1, 1], ship_lt=[2, 2, 2, 2], holding_cost=[0.5] * 4, backorder_cost=[1] * 4,
                 coplayer_target=[19, 20, 14, 25], arc_order=[3, 2, 1, 0], agents=(0, 2))   # shop, [dist_a], dist_b
    assert c is not None


@pytest.mark.skipif(not __import__("oracle.ref_harness", fromlist=["x"]).reference_available(),
                    reason="/root/reference not mounted (GPU box)")
def test_tree_oracle_vs_live_reference():
    from oracle import ref_harness as H
    for cid in (0, 17, 33, 58, 71, 90):
        c = TREE_CASES[cid]
        names = c["names"]
        nd = [dict(name=names[k], target=c["target"][k], holding=c["holding"][k], backorder=c["backorder"][k],
                   demand=tuple(c["demand"][k]) if c["demand"][k] else None) for k in range(len(names))]
        arcs = [("external_supplier" if c["supplier"][k] < 0 else names[c["supplier"][k]], names[k], c["li"][k], c["ls"][k])
                for k in c["arc_order"]]
        env = H.make_network_env(nd, arcs, [names[a] for a in c["agents"]], c["glob"],
                                 dplusa=(-8, 0, 16) if c["rule"] == "d+a" else None)
        acts = np.random.RandomState(cid).randint(0, 17, (100, len(c["agents"])))
        r = H.network_rollout(env, names, [a for a in acts], seed=5000 + cid)
        inv0, so, sh, dem = r["init"]
        init = np.zeros(36, dtype=np.int64)
        init[:len(inv0)] = inv0
        for k in range(len(inv0)):
            init[4 + 4 * k: 4 + 4 * k + len(so[k])] = so[k]
            init[20 + 4 * k: 20 + 4 * k + len(sh[k])] = sh[k]
        o0, obs, rew = tree_oracle_rollout(c, init[None], np.asarray(dem)[None], acts[:, None, :])
        assert np.array_equal(o0[0], r["obs0"]) and np.array_equal(obs[:, 0], r["obs"]) and np.array_equal(rew[:, 0], r["rew"])


# ---------------------------------------------------------------------- host side of the product on trees
def tree_product_spec(case, t_max=100):
    """Product ScenarioSpec of a tree golden case (multi_echelon_drl_b200.spec.tree_spec)."""
    from multi_echelon_drl_b200.spec import tree_spec
    names = case["names"]
    nodes = [dict(name=names[k], holding_cost=case["holding"][k], backorder_cost=case["backorder"][k], target=case["target"][k],
                  demand=tuple(case["demand"][k]) if case["demand"][k] else None) for k in range(len(names))]
    arcs = [("external_supplier" if case["supplier"][k] < 0 else names[case["supplier"][k]], names[k], case["li"][k], case["ls"][k])
            for k in case["arc_order"]]
    sp = tree_spec(nodes, arcs, [names[a] for a in case["agents"]], case["glob"], t_max)
    return sp.with_rule("d+a", -8, 0, 16) if case["rule"] == "d+a" else sp


def test_product_tree_specs_order_sequence_streams_and_c_validation():
    """The host side agrees with the reference on every golden tree: order sequence (utils/graph.py:5-32), the RNG
    draw order of SupplyNetwork.reset (initial state and demand streams from one seed), and the C library accepts
    the lowered spec with the right buffer sizes."""
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs, seeded_episode_inputs
    lib = _lib.load()
    for c in TREE_CASES + BIG_TREE_CASES:
        sp = tree_product_spec(c)
        big = len(c["names"]) > 4
        assert list(sp.order_sequence) == c["order_sequence"], c["id"]
        init, dem = seeded_episode_inputs(sp, [c["seed"]])
        assert np.array_equal(init[0], c["init"]) and np.array_equal(dem[0], c["demand_table"]), c["id"]
        cs = sp.to_c()
        assert lib.sn_spec_validate(C.byref(cs), _lib.SN_I32) == 0, lib.sn_last_error()
        assert lib.sn_state_words(C.byref(cs)) == (128 if big else 64) and lib.sn_init_words(C.byref(cs)) == (72 if big else 36)
        assert lib.sn_obs_dim(C.byref(cs)) == c["obs0"].shape[0] == sp.obs_dim
        binit, bdem = bulk_episode_inputs(sp, 16, np.random.RandomState(0))
        assert binit.shape == (16, 72 if big else 36) and bdem.shape == (16, sp.n_sources, 103)
        assert not binit[:, len(c["names"]):init_layout(len(c["names"]))[0]].any()
        for s, k in enumerate(sp.source_nodes):
            assert bdem[:, s].max() <= c["demand"][k][1]


def test_tree_spec_and_c_validation_errors():
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    from multi_echelon_drl_b200.spec import tree_spec
    lib = _lib.load()
    nodes = [dict(name="shop", demand=(0, 8)), dict(name="dist_a"), dict(name="dist_b", demand=(0, 5)), dict(name="maker")]
    arcs = [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)]
    sp = tree_spec(nodes, arcs, ["dist_a", "dist_b"])
    assert sp.order_sequence == ("shop", "dist_a", "dist_b", "maker") and sp.customer_rank == (0, 1, 0, 0)
    assert sp.n_sources == 2 and sp.source_nodes == (0, 2) and sp.external_name == "ext"
    with pytest.raises(ValueError):                     # shop, [dist_a], dist_b: not contiguous in the order sequence
        tree_spec(nodes, arcs, ["shop", "dist_b"])
    with pytest.raises(NotImplementedError):            # two suppliers for one node
        tree_spec(nodes, arcs + [("dist_b", "shop", 1, 1)], ["shop"])
    with pytest.raises(ValueError):                     # a node without an upstream arc
        tree_spec(nodes, arcs[:3], ["maker"])
    with pytest.raises(ValueError):                     # no demand source
        tree_spec([dict(name="a"), dict(name="b")], [("ext", "b", 1, 1), ("b", "a", 1, 1)], ["a"])
    with pytest.raises(ValueError):                     # cycle
        tree_spec([dict(name="a", demand=(0, 4)), dict(name="b")], [("a", "b", 1, 1), ("b", "a", 1, 1)], ["a"])
    c = sp.to_c()
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) == 0
    c.order_pos[1], c.order_pos[3] = c.order_pos[3], c.order_pos[1]          # supplier would order before its customer
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) != 0 and b"order" in lib.sn_last_error()
    c = sp.to_c()
    c.customer_rank[1] = 0                                                    # collides with dist_b's rank
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) != 0 and b"rank" in lib.sn_last_error()
    c = sp.to_c()
    c.supplier[0] = 0
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) != 0
    c = sp.to_c()
    c.n_items = 2
    assert lib.sn_spec_validate(C.byref(c), _lib.SN_I32) != 0


def test_yaml_description_of_a_tree_gives_the_golden_spec():
    """network_configs/distribution_tree.yaml (the reference's YAML schema) lowers to the same spec as the node / arc
    lists of the 'two_level' golden topology."""
    import dataclasses
    import multi_echelon_drl_b200 as pkg
    from multi_echelon_drl_b200.spec import specs_from_config
    from multi_echelon_drl_b200.utils.graph import read_yaml
    cfg = read_yaml(os.path.join(os.path.dirname(pkg.__file__), "network_configs", "distribution_tree.yaml"))
    c = [c for c in TREE_CASES if c["topo"] == "two_level" and c["agents"] == [1, 2] and c["glob"] and c["rule"] == "a"][0]
    (sp,) = specs_from_config(cfg, ["dist_a", "dist_b"], True)
    a, b = dataclasses.asdict(sp), dataclasses.asdict(tree_product_spec(c))
    a.pop("name"); b.pop("name")
    assert a == b
    with pytest.raises(NotImplementedError):              # a node with two suppliers is not a distribution tree
        bad = dict(cfg, arcs=cfg["arcs"] + [dict(customer="shop", supplier="dist_b", info_leadtime=1, shipment_leadtime=1)])
        specs_from_config(bad, ["shop"], True)


# ---------------------------------------------------------------------- random trees
def random_tree_case(rs, n_nodes=None):
    """A random distribution tree of 2..4 internal nodes: random supplier structure, names (they drive the order
    sequence), arc-list order (priority among customers), lead times with info + shipment <= 5, demand sources at
    the leaves and sometimes elsewhere, and an agent set that is contiguous in the resulting order sequence."""
    from oracle.network_np import order_sequence
    N = int(n_nodes or rs.randint(2, 5))           # pass n_nodes for the larger ones (5..8)
    pool = ["alpha", "bravo", "charlie", "delta", "echo", "foxtrot", "golf", "hotel"]
    names = list(rs.choice(pool, N, replace=False))
    supplier = [EXTERNAL] * N
    order = list(rs.permutation(N))               # order[i] may only buy from order[j], j < i, or the external supplier
    for i in range(1, N):
        if rs.rand() < 0.8:
            supplier[order[i]] = int(order[rs.randint(0, i)])
    has_customer = [any(supplier[c] == k for c in range(N)) for k in range(N)]
    demand = [((0, int(rs.randint(3, 10))) if (not has_customer[k] or rs.rand() < 0.25) else None) for k in range(N)]
    li, ls = [], []
    for _ in range(N):
        a = int(rs.randint(1, 5)); b = int(rs.randint(1, 6 - a)) if a < 5 else 1
        li.append(a); ls.append(min(b, 4))
    arc_order = [int(x) for x in rs.permutation(N)]
    cd = {names[k]: [names[c] for c in arc_order if supplier[c] == k] for k in range(N)}
    cd["external_supplier"] = [names[c] for c in arc_order if supplier[c] == EXTERNAL]
    seq = [s for s in order_sequence(names + ["external_supplier"], cd) if s != "external_supplier"]
    m = int(rs.randint(1, N + 1)); start = int(rs.randint(0, N - m + 1))
    agents = [names.index(s) for s in seq[start:start + m]]
    rs.shuffle(agents)
    return dict(names=names, supplier=supplier, demand=demand, li=li, ls=ls, holding=[float(rs.choice([0.25, 0.5, 1.0])) for _ in range(N)],
                backorder=[float(rs.choice([0.5, 1.0, 2.0])) for _ in range(N)], target=[int(rs.randint(8, 40)) for _ in range(N)],
                arc_order=arc_order, agents=[int(a) for a in agents], glob=bool(rs.rand() < 0.5), rule=str(rs.choice(["a", "d+a"])),
                order_sequence=seq)


def random_tree_inputs(case, n, rs):
    N = len(case["names"])
    _, so0, sh0 = init_layout(N)
    init = np.zeros((n, 72 if N > 4 else 36), dtype=np.int64)
    init[:, :N] = rs.randint(0, 25, (n, N))
    for k in range(N):
        init[:, so0 + 4 * k: so0 + 4 * k + case["li"][k]] = rs.randint(0, 9, (n, case["li"][k]))
        init[:, sh0 + 4 * k: sh0 + 4 * k + case["ls"][k]] = rs.randint(0, 9, (n, case["ls"][k]))
    dem = np.stack([rs.randint(d[0], d[1] + 1, (n, 103)) for d in case["demand"] if d is not None], axis=1)
    acts = rs.randint(-1, 19, (100, n, len(case["agents"])))
    return init, dem, acts


@pytest.mark.skipif(not __import__("oracle.ref_harness", fromlist=["x"]).reference_available(),
                    reason="/root/reference not mounted (GPU box)")
def test_random_trees_vs_live_reference():
    """56 random trees (16 of them with five to eight nodes): the restatement and the reference's own classes agree on every observation, reward and
    per-node cost, and on the order sequence."""
    from oracle import ref_harness as H
    rs = np.random.RandomState(2024)
    for i in range(56):
        c = random_tree_case(rs, None if i < 40 else 5 + i % 4)          # the last 16: five to eight nodes
        names = c["names"]
        nd = [dict(name=names[k], target=c["target"][k], holding=c["holding"][k], backorder=c["backorder"][k],
                   demand=c["demand"][k]) for k in range(len(names))]
        arcs = [("external_supplier" if c["supplier"][k] < 0 else names[c["supplier"][k]], names[k], c["li"][k], c["ls"][k])
                for k in c["arc_order"]]
        env = H.make_network_env(nd, arcs, [names[a] for a in c["agents"]], c["glob"],
                                 dplusa=(-8, 0, 16) if c["rule"] == "d+a" else None)
        acts = rs.randint(-1, 19, (100, len(c["agents"])))
        r = H.network_rollout(env, names, [a for a in acts], seed=7000 + i)
        assert [s for s in r["order_sequence"] if s != "external_supplier"] == c["order_sequence"], (i, c)
        inv0, so, sh, dem = r["init"]
        _, so0, sh0 = init_layout(len(inv0))
        init = np.zeros(72 if len(inv0) > 4 else 36, dtype=np.int64)
        init[:len(inv0)] = inv0
        for k in range(len(inv0)):
            init[so0 + 4 * k: so0 + 4 * k + len(so[k])] = so[k]
            init[sh0 + 4 * k: sh0 + 4 * k + len(sh[k])] = sh[k]
        o0, obs, rew, hold, back = tree_oracle_rollout(c, init[None], np.asarray(dem)[None], acts[:, None, :], return_costs=True)
        assert np.array_equal(o0[0], r["obs0"]), (i, c)
        assert np.array_equal(obs[:, 0], r["obs"]) and np.array_equal(rew[:, 0], r["rew"]), (i, c)
        assert np.array_equal(hold[:, 0], r["holding"]) and np.array_equal(back[:, 0], r["backorder"]), (i, c)


def test_random_tree_specs_lower_consistently():
    """Product spec of random trees: same order sequence as the oracle's restatement of utils/graph.py, accepted by
    the C validation; quantities never go negative in the oracle (inventory, backlog, pipelines)."""
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(77)
    for i in range(80):
        c = random_tree_case(rs, None if i < 60 else 5 + i % 4)
        sp = tree_product_spec(c)
        assert list(sp.order_sequence) == c["order_sequence"], c
        cs = sp.to_c()
        assert lib.sn_spec_validate(C.byref(cs), _lib.SN_I32) == 0, (lib.sn_last_error(), c)
        if i < 10:
            init, dem, acts = random_tree_inputs(c, 8, rs)
            orc = TreeOracle(oracle_spec(c), 8)
            N = len(c["names"])
            _, so0, sh0 = init_layout(N)
            orc.reset(init[:, :N], [init[:, so0 + 4 * k: so0 + 4 * k + c["li"][k]] for k in range(N)],
                      [init[:, sh0 + 4 * k: sh0 + 4 * k + c["ls"][k]] for k in range(N)], dem)
            for t in range(100):
                orc.step(acts[t])
                orc.check_invariant()
                assert all((x >= 0).all() for x in orc.inv + orc.B + orc.unf_ind)


def test_eight_node_spec_host_side():
    """Host logic for trees of five to eight nodes: 72-word init layout, per-source demand tables, 8-wide C arrays,
    limits (nine nodes, chains longer than four)."""
    import ctypes as C
    from dataclasses import replace
    from multi_echelon_drl_b200 import _lib, uniform_spec
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.utils.demands import _fixed_init, bulk_episode_inputs, init_layout as product_layout
    c = BIG_TREE_CASES[[x["topo"] for x in BIG_TREE_CASES].index("eight_hub")]
    sp = tree_product_spec(c)
    assert sp.n_facilities == 8 and sp.n_sources == 6 and sp.obs_dim == 7 * (8 if c["glob"] else len(c["agents"]))
    n_words, so_off, sh_off = product_layout(sp)
    assert n_words == 72 and so_off[:2] == (8, 12) and sh_off[:2] == (40, 44)
    init, dem = bulk_episode_inputs(sp, 32, np.random.RandomState(0))
    assert init.shape == (32, 72) and dem.shape == (32, 6, 103)
    for k in range(8):                                     # slots beyond an arc's lead time stay empty
        assert not init[:, so_off[k] + sp.info_leadtime[k]: so_off[k] + 4].any()
        assert not init[:, sh_off[k] + sp.ship_leadtime[k]: sh_off[k] + 4].any()
    fixed = _fixed_init(replace(sp, random_init=False))
    assert fixed.shape == (72,) and (fixed[:8] == sp.fixed_inventory).all()
    cs = sp.to_c()
    assert len(cs.supplier) == 8 and list(cs.supplier) == list(sp.suppliers) and cs.topology == _lib.SN_TOPO_TREE
    lib = _lib.load()
    assert lib.sn_spec_validate(C.byref(cs), _lib.SN_F32) == 0
    cs.n_facilities = 9
    assert lib.sn_spec_validate(C.byref(cs), _lib.SN_F32) == _lib.SN_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):               # nine internal nodes
        tree_spec([dict(name=f"n{i}", demand=(0, 4) if i == 0 else None) for i in range(9)],
                  [("ext", "n8", 1, 1)] + [(f"n{i + 1}", f"n{i}", 1, 1) for i in range(8)], ["n0"])
    with pytest.raises(ValueError):                        # a plain chain spec stays limited to four facilities
        replace(uniform_spec(None, True), n_facilities=5)
    chain5 = uniform_spec(None, True).to_c()               # and so does the C side
    chain5.n_facilities = 5
    assert lib.sn_spec_validate(C.byref(chain5), _lib.SN_I32) == _lib.SN_ERR_UNSUPPORTED


@pytest.mark.skipif(not __import__("oracle.ref_harness", fromlist=["x"]).reference_available(),
                    reason="/root/reference not mounted (GPU box)")
def test_tree_oracle_float_actions_vs_live_reference():
    """Fractional (TD3-like) float32 actions on trees: the reference runs them in mixed float32 / Python arithmetic
    (numpy >= 2 scalar rules); the float64 restatement agrees to 1e-6 relative plus an absolute term for cancellation
    residues of O(100) quantities - the tolerance the GPU float tests use against the restatement."""
    from oracle import ref_harness as H
    rs = np.random.RandomState(515)
    for i in range(12):
        c = random_tree_case(rs, None if i < 8 else 5 + i % 4)
        names = c["names"]
        nd = [dict(name=names[k], target=c["target"][k], holding=c["holding"][k], backorder=c["backorder"][k],
                   demand=c["demand"][k]) for k in range(len(names))]
        arcs = [("external_supplier" if c["supplier"][k] < 0 else names[c["supplier"][k]], names[k], c["li"][k], c["ls"][k])
                for k in c["arc_order"]]
        env = H.make_network_env(nd, arcs, [names[a] for a in c["agents"]], c["glob"],
                                 dplusa=(-8, 0, 16) if c["rule"] == "d+a" else None)
        acts = (rs.rand(100, len(c["agents"])) * 18 - 1).astype(np.float32)
        r = H.network_rollout(env, names, [a for a in acts], seed=8000 + i)
        inv0, so, sh, dem = r["init"]
        N = len(inv0)
        orc = TreeOracle(oracle_spec(c), 1, dtype=np.float64)
        o0 = orc.reset(np.array([inv0], dtype=np.float64), [np.array([x], dtype=np.float64) for x in so],
                       [np.array([x], dtype=np.float64) for x in sh], np.asarray(dem)[None])
        assert np.array_equal(o0[0], r["obs0"])
        for t in range(100):
            o, rew, _ = orc.step(acts[t][None].astype(np.float64))
            np.testing.assert_allclose(o[0], r["obs"][t], rtol=1e-6, atol=2e-4, err_msg=str((i, t)))
            np.testing.assert_allclose(rew[0], r["rew"][t], rtol=1e-6, atol=2e-3, err_msg=str((i, t)))
        assert N == len(names)
