"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference is pure Python and cannot travel to the GPU box, so its outputs are committed
here as small fixtures; every test that needs "what the reference does" reads these files.

Files written
  rollouts.npz    per case: inputs (init state, demand, actions) and the reference's outputs
                  (obs0, obs[T], reward[T], done[T], per-facility holding/backorder cost[T,4])
  heuristic.npz   BaseStockPolicy.get_order_quantity (utils/heuristics.py:35-75) on single-row
                  observations, ndarray path and dict path, rules 'a' and 'd+a'
  seeded.npz      `np.random.seed(s); env.reset()` initial state + demand for both scenarios
                  (pins the RNG draw order replayed by the host-side stream generator)
  classic.npz     the classic beer game episodes behind the 13 totals of
                  tests/test_supplynetwork.py:25-69, with per-step outputs
  leadtimes.npz   serial chains with other per-arc lead times (general-chain path)
  chains.npz      serial chains with 1..3 facilities
  trees.npz       distribution trees (one supplier per node, several customers and / or own demand per node)
  trees8.npz      the same with five to eight internal nodes (72-word initial-state layout)
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NODE = H.NODE_NAMES
SCN = {"uniform": dict(hi=16, dpa=(-8, 0, 16)), "normal": dict(hi=20, dpa=(-10, 0, 20))}


def flat_init(init):
    inv0, so, sh, dem = init
    # 19 numbers in the product's init order: inv[4], so[k][*] (k=0..3), sh[k][*] (k=0..3)
    v = list(inv0)
    for k in range(4):
        v += list(so[k])
    for k in range(4):
        v += list(sh[k])
    return np.array(v, dtype=np.int64), np.asarray(dem, dtype=np.int64)


def gen_rollouts():
    cases = {}
    meta = []
    agsets = [(0, 1, 2, 3), (0,), (1,), (2,), (3,), (0, 1), (1, 2, 3), (2, 3), (3, 2, 1, 0)]
    cid = 0
    for scn, par in SCN.items():
        for ag in agsets:
            for glob in (False, True):
                for rule in ("a", "d+a"):
                    for kind in ("int", "float"):
                        if kind == "float" and (glob or len(ag) not in (1, 4)):
                            continue
                        seed = 100 + cid
                        T = 100
                        env = H.make_env(scn, [NODE[k] for k in ag], glob, True, T,
                                         dplusa=par["dpa"] if rule == "d+a" else None)
                        rs = np.random.RandomState(5000 + cid)
                        if kind == "int":
                            lo = -2 if rule == "a" else 0           # negatives exercise the clamp
                            acts = rs.randint(lo, par["hi"] + 4, (T, len(ag))).astype(np.int64)
                            feed = [a for a in acts]
                        else:
                            acts = (rs.rand(T, len(ag)) * par["hi"]).astype(np.float32)
                            feed = [a for a in acts]
                        r = H.rollout(env, feed, seed=seed, collect_costs=True)
                        init, dem = flat_init(r["init"])
                        p = f"c{cid:03d}_"
                        cases[p + "init"] = init
                        cases[p + "demand"] = dem
                        cases[p + "actions"] = acts
                        cases[p + "obs0"] = r["obs0"]
                        cases[p + "obs"] = r["obs"]
                        cases[p + "rew"] = r["rew"]
                        cases[p + "done"] = r["done"]
                        cases[p + "holding"] = r["holding"]
                        cases[p + "backorder"] = r["backorder"]
                        meta.append(f"{cid}|{scn}|{','.join(map(str, ag))}|{int(glob)}|{rule}|{kind}|{seed}")
                        cid += 1
    cases["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "rollouts.npz"), **cases)
    print("rollouts:", cid, "cases")


def gen_heuristic():
    ref = H.import_reference()
    BSP = ref["heuristics"].BaseStockPolicy
    idx = {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1, "latest_demand": 2}
    rs = np.random.RandomState(7)
    out = {}
    n = 64
    for name, targets, lb, ub in (("uniform", [19, 20, 20, 14], 0, 16), ("normal", [48, 43, 41, 30], 0, 20)):
        obs = rs.randint(0, 30, (n, 28)).astype(np.float32)
        obs[n // 2:] += rs.rand(n - n // 2, 28).astype(np.float32)      # fractional rows
        for rule in ("a", "d+a"):
            pol = BSP(targets, idx, 7, lb, ub, rule)
            nd = np.stack([pol.get_order_quantity(obs[i:i + 1]) for i in range(n)])     # ndarray path
            keys = ["on_hand", "unfilled_demand", "latest_demand"] + [f"unreceived_pipeline_{i}" for i in range(4)]
            dct = np.zeros((n, 4))
            for i in range(n):
                for f in range(4):
                    p1 = BSP([targets[f]], idx, 7, lb, ub, rule)
                    d = {k: obs[i, 7 * f + j] for j, k in enumerate(keys)}
                    dct[i, f] = p1.get_order_quantity({"x": d}).item()               # dict path
            out[f"{name}_{rule}_nd"] = nd
            out[f"{name}_{rule}_dict"] = dct
        out[f"{name}_obs"] = obs
        # the reference's batch quirk (row 0 only), SURVEY Q7
        pol = BSP(targets, idx, 7, lb, ub, "a")
        out[f"{name}_row0_batch2"] = pol.get_order_quantity(obs[:2])
    np.savez_compressed(os.path.join(OUT, "heuristic.npz"), **out)
    print("heuristic ok")


def gen_seeded():
    out = {}
    for scn in SCN:
        env = H.make_env(scn, NODE, True, True, 100)
        for seed in (0, 1, 2, 12345, 2**31 - 1):
            obs0, init = H.reset_and_capture(env, seed)
            i, d = flat_init(init)
            out[f"{scn}_{seed}_init"] = i
            out[f"{scn}_{seed}_demand"] = d
            out[f"{scn}_{seed}_obs0"] = obs0
    np.savez_compressed(os.path.join(OUT, "seeded.npz"), **out)
    print("seeded ok")


def gen_classic():
    """tests/test_supplynetwork.py:25-69 — the totals themselves live in the test file."""
    out = {}
    z, f, w = [0] * 20, [4] * 20, [12] * 20
    e = [0, 0] + [8] * 18
    runs = [("retailer", (0,), [z, f, w]), ("manufacturer", (3,), [e, f, w]),
            ("distributor", (2,), [e, f, w]), ("wholesaler", (1,), [e, f, w])]
    cid = 0
    for name, ag, qs in runs:
        env = H.make_env("classic", [name], False, False, 35)
        for q in qs:
            acts = np.array(q).reshape(20, 1)
            r = H.rollout(env, [a for a in acts], collect_costs=True)
            i, d = flat_init(r["init"])
            p = f"k{cid:02d}_"
            out.update({p + "init": i, p + "demand": d, p + "actions": acts, p + "obs0": r["obs0"],
                        p + "obs": r["obs"], p + "rew": r["rew"], p + "agents": np.array(ag)})
            cid += 1
    env = H.make_env("classic", ["retailer", "wholesaler"], False, False, 35)
    acts = np.array([0, 8] + [0, 4] + [0, 0] * 18).reshape(20, 2)
    r = H.rollout(env, [a for a in acts], collect_costs=True)
    i, d = flat_init(r["init"])
    p = f"k{cid:02d}_"
    out.update({p + "init": i, p + "demand": d, p + "actions": acts, p + "obs0": r["obs0"],
                p + "obs": r["obs"], p + "rew": r["rew"], p + "agents": np.array([0, 1])})
    np.savez_compressed(os.path.join(OUT, "classic.npz"), **out)
    print("classic:", cid + 1, "episodes; totals:",
          [float(out[f"k{c:02d}_rew"].sum()) for c in range(cid + 1)])


def flat_init_generic(init):
    """36 numbers: inv[4], so[k][0..3] (k = 0..3, unused slots 0), sh[k][0..3]; shorter chains leave the
    rows of the missing facilities at 0."""
    inv0, so, sh, dem = init
    v = np.zeros(36, dtype=np.int64)
    v[:len(inv0)] = inv0
    for k in range(len(inv0)):
        v[4 + 4 * k: 4 + 4 * k + len(so[k])] = so[k]
        v[20 + 4 * k: 20 + 4 * k + len(sh[k])] = sh[k]
    return v, np.asarray(dem, dtype=np.int64)


def gen_leadtimes():
    """Serial chains with per-arc lead times other than the beer game's, built from the reference's own
    Node/Arc/SupplyNetwork classes (oracle/ref_harness.make_custom_chain_env).  info+shipment <= 5 per arc:
    beyond that the reference's own on_order assertion (supplynetwork.py:159-160) fails at reset."""
    combos = [((2, 2, 2, 1), (2, 2, 2, 2)), ((1, 1, 1, 1), (1, 1, 1, 1)), ((3, 2, 1, 2), (2, 3, 4, 1)),
              ((4, 1, 2, 3), (1, 4, 3, 2)), ((1, 4, 1, 2), (4, 1, 4, 3)), ((2, 3, 2, 2), (3, 2, 1, 3))]
    out, meta, cid = {}, [], 0
    for li, ls in combos:
        for ag in ((0, 1, 2, 3), (1,), (2, 3), (3, 2, 1, 0)):
            for glob, rule in ((False, "a"), (True, "d+a"), (True, "a")):
                seed = 900 + cid
                env = H.make_custom_chain_env(li, ls, ag, glob, dplusa=(-8, 0, 16) if rule == "d+a" else None)
                acts = np.random.RandomState(7000 + cid).randint(-1, 19, (100, len(ag))).astype(np.int64)
                r = H.rollout(env, [a for a in acts], seed=seed, collect_costs=True)
                init, dem = flat_init_generic(r["init"])
                p = f"g{cid:03d}_"
                out.update({p + "init": init, p + "demand": dem, p + "actions": acts, p + "obs0": r["obs0"],
                            p + "obs": r["obs"], p + "rew": r["rew"]})
                meta.append(f"{cid}|{','.join(map(str, li))}|{','.join(map(str, ls))}|{','.join(map(str, ag))}|{int(glob)}|{rule}")
                cid += 1
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "leadtimes.npz"), **out)
    print("leadtimes:", cid, "cases")


def gen_short_chains():
    """Serial chains with fewer than four facilities (first F of retailer..manufacturer, then the external
    supplier), again built from the reference's own classes."""
    cases = [((2, 1), (2, 2), [(0, 1), (0,), (1,), (1, 0)]), ((1,), (2,), [(0,)]),
             ((2, 2, 1), (2, 2, 2), [(0, 1, 2), (1,), (1, 2), (0,)]), ((3, 1, 2), (1, 4, 3), [(0, 1, 2), (2,), (2, 1, 0)]),
             ((1, 3), (4, 2), [(0, 1), (1,)])]
    out, meta, cid = {}, [], 0
    for li, ls, agsets in cases:
        nf = len(li)
        for ag in agsets:
            for glob, rule in ((False, "a"), (True, "d+a"), (True, "a")):
                seed = 1300 + cid
                env = H.make_custom_chain_env(li, ls, ag, glob, targets=(19, 20, 20, 14)[:nf],
                                              dplusa=(-8, 0, 16) if rule == "d+a" else None)
                acts = np.random.RandomState(8000 + cid).randint(-1, 19, (100, len(ag))).astype(np.int64)
                r = H.rollout(env, [a for a in acts], seed=seed, collect_costs=True)
                init, dem = flat_init_generic(r["init"])
                p = f"s{cid:03d}_"
                out.update({p + "init": init, p + "demand": dem, p + "actions": acts, p + "obs0": r["obs0"],
                            p + "obs": r["obs"], p + "rew": r["rew"]})
                meta.append(f"{cid}|{','.join(map(str, li))}|{','.join(map(str, ls))}|{','.join(map(str, ag))}|{int(glob)}|{rule}")
                cid += 1
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "chains.npz"), **out)
    print("short chains:", cid, "cases")


# Distribution trees.  Per topology: nodes in node-list order as (name, supplier name | "EXT", demand range |
# None, base-stock target, holding cost, backorder cost, info lead time, shipment lead time of its upstream arc)
# and the arc-list order (arcs named by their customer).  Names and arc order are chosen so that the reference's
# order sequence (utils/graph.py:5-32) differs from the node-list order in some of them.
TREES = {
    "two_ret": ([("ret_a", "whole", (0, 8), 19, 0.5, 1, 2, 2), ("ret_b", "whole", (0, 4), 15, 0.5, 1, 1, 3),
                 ("whole", "EXT", None, 30, 0.5, 1, 1, 2)], ["whole", "ret_a", "ret_b"]),
    "two_ret_rev": ([("whole", "EXT", None, 30, 0.25, 2, 2, 2), ("ret_b", "whole", (0, 4), 15, 0.5, 1, 1, 3),
                     ("ret_a", "whole", (0, 8), 19, 1.0, 0.5, 2, 2)], ["ret_b", "ret_a", "whole"]),
    "three_ret": ([("r1", "hub", (0, 8), 19, 0.5, 1, 2, 2), ("r2", "hub", (0, 6), 12, 0.5, 1, 2, 1),
                   ("r3", "hub", (0, 3), 9, 0.5, 1, 1, 1), ("hub", "EXT", None, 40, 0.5, 1, 2, 3)],
                  ["hub", "r3", "r1", "r2"]),
    "two_level": ([("shop", "dist_a", (0, 8), 19, 0.5, 1, 2, 2), ("dist_a", "maker", None, 20, 0.5, 1, 2, 2),
                   ("dist_b", "maker", (0, 5), 14, 0.5, 1, 1, 2), ("maker", "EXT", None, 25, 0.5, 1, 1, 2)],
                  ["maker", "dist_b", "dist_a", "shop"]),
    "src_with_cust": ([("outlet", "store", (0, 5), 12, 0.5, 1, 1, 2), ("store", "depot", (0, 6), 22, 0.5, 1, 2, 2),
                       ("depot", "EXT", None, 25, 0.5, 1, 1, 3)], ["depot", "store", "outlet"]),
    "two_roots": ([("a_shop", "a_dc", (0, 8), 19, 0.5, 1, 2, 2), ("a_dc", "EXT", None, 20, 0.5, 1, 1, 2),
                   ("b_shop", "b_dc", (0, 8), 19, 0.5, 1, 2, 2), ("b_dc", "EXT", None, 20, 0.5, 1, 2, 1)],
                  ["a_dc", "a_shop", "b_dc", "b_shop"]),
    "chain_as_tree": ([("retailer", "wholesaler", (0, 8), 19, 0.5, 1, 2, 2), ("wholesaler", "distributor", None, 20, 0.5, 1, 2, 2),
                       ("distributor", "manufacturer", None, 20, 0.5, 1, 2, 2), ("manufacturer", "EXT", None, 14, 0.5, 1, 1, 2)],
                      ["manufacturer", "distributor", "wholesaler", "retailer"]),
}
TREE_AGENTS = {   # node indices, caller order; all contiguous in the respective order sequence
    "two_ret": [(0, 1, 2), (2,), (1,), (1, 2), (2, 1, 0)],
    "two_ret_rev": [(0, 1, 2), (0,), (2,), (2, 0), (1, 2)],
    "three_ret": [(0, 1, 2, 3), (3,), (2,), (1, 3), (0, 1), (3, 1, 0)],
    "two_level": [(0, 1, 2, 3), (3,), (1,), (2, 3), (1, 2), (0, 1, 2)],
    "src_with_cust": [(0, 1, 2), (1,), (2,), (0, 1)],
    "two_roots": [(0, 1, 2, 3), (1, 2), (3,), (0,), (2, 3)],
    "chain_as_tree": [(0, 1, 2, 3), (2,), (3, 2)],
}


BIG_TREES = {
    "seven_two_levels": ([("s1", "dc_east", (0, 8), 19, 0.5, 1, 2, 2), ("s2", "dc_east", (0, 5), 13, 0.5, 1, 1, 2),
                          ("s3", "dc_west", (0, 6), 15, 0.5, 1, 2, 1), ("s4", "dc_west", (0, 3), 9, 0.5, 2, 1, 1),
                          ("dc_east", "plant", None, 30, 0.25, 1, 2, 2), ("dc_west", "plant", None, 24, 0.25, 1, 1, 3),
                          ("plant", "EXT", None, 50, 0.125, 1, 1, 2)],
                         ["plant", "dc_west", "dc_east", "s4", "s1", "s3", "s2"]),
    "eight_hub": ([("hub", "factory", None, 60, 0.25, 1, 2, 2), ("r1", "hub", (0, 8), 19, 0.5, 1, 2, 2), ("r2", "hub", (0, 6), 14, 0.5, 1, 1, 1),
                   ("r3", "hub", (0, 4), 10, 0.5, 1, 2, 1), ("r4", "hub", (0, 3), 8, 0.5, 1, 1, 2), ("r5", "hub", (0, 2), 6, 1.0, 1, 3, 1),
                   ("factory", "EXT", None, 40, 0.125, 1, 1, 2), ("annex", "factory", (0, 5), 12, 0.5, 1, 1, 3)],
                  ["factory", "annex", "hub", "r3", "r1", "r5", "r2", "r4"]),
    "five_mixed": ([("kiosk", "shop_b", (0, 3), 8, 0.5, 1, 1, 1), ("shop_a", "dist", (0, 8), 19, 0.5, 1, 2, 2),
                    ("shop_b", "dist", (0, 5), 16, 0.5, 1, 2, 2), ("dist", "maker", None, 35, 0.5, 1, 2, 2),
                    ("maker", "EXT", None, 30, 0.5, 1, 1, 2)], ["maker", "dist", "shop_b", "shop_a", "kiosk"]),
    "six_two_roots": ([("a1", "a_dc", (0, 8), 19, 0.5, 1, 2, 2), ("a2", "a_dc", (0, 4), 10, 0.5, 1, 1, 2), ("a_dc", "EXT", None, 30, 0.5, 1, 1, 2),
                       ("b1", "b_dc", (0, 6), 15, 0.5, 1, 2, 1), ("b_dc", "b_plant", None, 20, 0.5, 1, 2, 2),
                       ("b_plant", "EXT", None, 18, 0.5, 1, 1, 3)], ["a_dc", "a2", "a1", "b_plant", "b_dc", "b1"]),
    "eight_chain": ([("n1", "n2", (0, 8), 19, 0.5, 1, 2, 2), ("n2", "n3", None, 20, 0.5, 1, 2, 2), ("n3", "n4", None, 20, 0.5, 1, 2, 2),
                     ("n4", "n5", None, 20, 0.5, 1, 1, 2), ("n5", "n6", None, 20, 0.5, 1, 2, 1), ("n6", "n7", None, 20, 0.5, 1, 1, 1),
                     ("n7", "n8", None, 20, 0.5, 1, 2, 2), ("n8", "EXT", None, 14, 0.5, 1, 1, 2)],
                    ["n8", "n7", "n6", "n5", "n4", "n3", "n2", "n1"]),
}


def flat_init_net(init):
    """72 numbers: inv[8]; so[k][0..3] at 8+4k; sh[k][0..3] at 40+4k (layout of the kernels for trees of 5..8 nodes)."""
    inv0, so, sh, _ = init
    v = np.zeros(72, dtype=np.int64)
    v[:len(inv0)] = inv0
    for k in range(len(inv0)):
        v[8 + 4 * k: 8 + 4 * k + len(so[k])] = so[k]
        v[40 + 4 * k: 40 + 4 * k + len(sh[k])] = sh[k]
    return v


def gen_big_trees():
    """Trees of five to eight internal nodes; agent sets are drawn as contiguous windows of the order sequence."""
    import json
    rs = np.random.RandomState(99)
    out, meta, cid = {}, [], 0
    for topo, (nodes, arc_order) in BIG_TREES.items():
        names = [n[0] for n in nodes]
        byname = {n[0]: n for n in nodes}
        nd = [dict(name=n[0], target=n[3], holding=n[4], backorder=n[5], demand=n[2]) for n in nodes]
        arcs = [("external_supplier" if byname[c][1] == "EXT" else byname[c][1], c, byname[c][6], byname[c][7])
                for c in arc_order]
        probe = H.make_network_env(nd, arcs, [names[0]], False)
        seq = [s for s in probe.sn.order_sequence if s != "external_supplier"]
        N = len(names)
        windows = [(0, N), (N - 1, 1), (0, 1), (1, 3), (N - 3, 3), (2, N - 2)]
        for (start, m) in windows:
            ag = [names.index(s) for s in seq[start:start + m]]
            if cid % 2:
                ag = ag[::-1]
            for glob, rule in ((False, "a"), (True, "d+a")) if (start, m) != (0, N) else ((True, "a"), (False, "d+a")):
                seed = 3100 + cid
                env = H.make_network_env(nd, arcs, [names[a] for a in ag], glob,
                                         dplusa=(-8, 0, 16) if rule == "d+a" else None)
                acts = rs.randint(-1, 19, (100, len(ag))).astype(np.int64)
                r = H.network_rollout(env, names, [a for a in acts], seed)
                inv0, so, sh, dem = r["init"]
                p = f"b{cid:03d}_"
                out.update({p + "init": flat_init_net(r["init"]), p + "demand": np.asarray(dem, dtype=np.int64), p + "actions": acts,
                            p + "obs0": r["obs0"], p + "obs": r["obs"], p + "rew": r["rew"],
                            p + "holding": r["holding"], p + "backorder": r["backorder"]})
                meta.append(json.dumps(dict(
                    cid=cid, topo=topo, names=names, supplier=[-1 if n[1] == "EXT" else names.index(n[1]) for n in nodes],
                    demand=[list(n[2]) if n[2] else None for n in nodes], target=[n[3] for n in nodes],
                    holding=[n[4] for n in nodes], backorder=[n[5] for n in nodes], li=[n[6] for n in nodes],
                    ls=[n[7] for n in nodes], arc_order=[names.index(c) for c in arc_order], agents=[int(a) for a in ag],
                    glob=glob, rule=rule, seed=seed, order_sequence=seq)))
                cid += 1
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "trees8.npz"), **out)
    print("trees of 5..8 nodes:", cid, "cases")


def gen_trees():
    import json
    out, meta, cid = {}, [], 0
    for topo, (nodes, arc_order) in TREES.items():
        names = [n[0] for n in nodes]
        byname = {n[0]: n for n in nodes}
        nd = [dict(name=n[0], target=n[3], holding=n[4], backorder=n[5], demand=n[2]) for n in nodes]
        arcs = [("external_supplier" if byname[c][1] == "EXT" else byname[c][1], c, byname[c][6], byname[c][7])
                for c in arc_order]
        for ag in TREE_AGENTS[topo]:
            for glob, rule in ((False, "a"), (True, "d+a"), (True, "a")):
                seed = 2100 + cid
                env = H.make_network_env(nd, arcs, [names[a] for a in ag], glob,
                                         dplusa=(-8, 0, 16) if rule == "d+a" else None)
                acts = np.random.RandomState(9000 + cid).randint(-1, 19, (100, len(ag))).astype(np.int64)
                r = H.network_rollout(env, names, [a for a in acts], seed)
                inv0, so, sh, dem = r["init"]
                init, _ = flat_init_generic((inv0, so, sh, []))
                p = f"t{cid:03d}_"
                out.update({p + "init": init, p + "demand": np.asarray(dem, dtype=np.int64), p + "actions": acts,
                            p + "obs0": r["obs0"], p + "obs": r["obs"], p + "rew": r["rew"],
                            p + "holding": r["holding"], p + "backorder": r["backorder"]})
                meta.append(json.dumps(dict(
                    cid=cid, topo=topo, names=names, supplier=[-1 if n[1] == "EXT" else names.index(n[1]) for n in nodes],
                    demand=[list(n[2]) if n[2] else None for n in nodes], target=[n[3] for n in nodes],
                    holding=[n[4] for n in nodes], backorder=[n[5] for n in nodes], li=[n[6] for n in nodes],
                    ls=[n[7] for n in nodes], arc_order=[names.index(c) for c in arc_order], agents=list(ag),
                    glob=glob, rule=rule, seed=seed,
                    order_sequence=[s for s in r["order_sequence"] if s != "external_supplier"])))
                cid += 1
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "trees.npz"), **out)
    print("trees:", cid, "cases")


if __name__ == "__main__":
    warnings.simplefilter("ignore", RuntimeWarning)
    gen_rollouts()
    gen_heuristic()
    gen_seeded()
    gen_classic()
    gen_leadtimes()
    gen_short_chains()
    gen_trees()
    gen_big_trees()
