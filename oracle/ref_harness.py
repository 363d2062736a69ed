"""Harness around the UNMODIFIED reference at /root/reference (TEST INFRASTRUCTURE ONLY).

This module only works in the build container, where /root/reference is mounted.  It is
used by `tests/golden/make_golden.py` (to generate the committed golden vectors) and by
the `-m "not gpu"` differential tests when the reference is present (they skip otherwise).
Nothing in the product package, `bench.py`, `smoke()` or the `-m gpu` tests imports it:
the GPU box has no /root/reference.

It does three things:
  * puts a stand-in `gym` package (oracle/gym_shim) on sys.path, because the reference's
    envs.py imports gym 0.21 which is not installed here;
  * builds reference envs (`make_beer_game*` factories, envs.py:163-596) and, optionally,
    applies the reference's d+a wrapper (utils/wrappers.py:6-31);
  * reads the post-reset initial state out of the reference objects so the same
    numbers can be fed to the oracle restatement and to the CUDA path.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("SUPPLYNET_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_shim")

NODE_NAMES = ["retailer", "wholesaler", "distributor", "manufacturer"]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "supplynetwork.py"))


def import_reference():
    """Import the reference modules in place (no copy) and return them as a namespace dict."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import envs as ref_envs  # noqa
    import supplynetwork as ref_sn  # noqa
    import network_components as ref_nc  # noqa
    from utils import heuristics as ref_heur, demands as ref_dem, wrappers as ref_wrap, graph as ref_graph  # noqa

    return dict(envs=ref_envs, sn=ref_sn, nc=ref_nc, heuristics=ref_heur, demands=ref_dem,
                wrappers=ref_wrap, graph=ref_graph)


def make_env(scenario: str, agents, global_observable=False, random_init=True, max_episode_steps=100,
             dplusa=None):
    """scenario in {'uniform','normal','classic'}; agents = list of node names (caller order).
    dplusa = None or (offset, lb, ub)."""
    ref = import_reference()
    E = ref["envs"]
    if scenario == "uniform":
        env = E.make_beer_game_uniform_multi_facility(
            agent_managed_facilities=list(agents), box_action_space=True, random_init=random_init,
            global_observable=global_observable, max_episode_steps=max_episode_steps)
    elif scenario == "normal":
        env = E.make_beer_game_normal_multi_facility(
            agent_managed_facilities=list(agents), box_action_space=True, random_init=random_init,
            global_observable=global_observable, max_episode_steps=max_episode_steps)
    elif scenario == "classic":
        env = E.make_beer_game(agent_managed_facilities=list(agents), max_period=max_episode_steps)
        env.global_observable = global_observable
    else:
        raise ValueError(scenario)
    if dplusa is not None:
        off, lb, ub = dplusa
        env = ref["wrappers"].wrap_action_d_plus_a(env, offset=off, lb=lb, ub=ub)
    return env


def read_initial_state(env):
    """Read (inv0[F], so[F][*], sh[F][*], demand[T+3]) out of a reference env right after
    `sn.reset()` and BEFORE `before_action(0)`.  Arc index k = arc whose customer is node k."""
    sn = env.sn
    names = [n for n in NODE_NAMES if n in sn.nodes]
    inv0 = [sn.nodes[n].current_inventory for n in names]
    sup = names[1:] + ["external_supplier"]
    so, sh = [], []
    for k, n in enumerate(names):
        arc = sn.arcs[(sup[k], n)]
        so.append([o.order_quantity for o in sorted(arc.sales_orders, key=lambda o: o.remaining_lead_time)])
        sh.append([s.quantity for s in sorted(arc.shipments, key=lambda s: s.time_till_arrival)])
    dem = np.asarray(sn.nodes["retailer"].demands._demands).copy()
    return inv0, so, sh, dem


def reset_and_capture(env, seed=None):
    """np.random.seed(seed); replicate env.reset() (envs.py:64-80) but capture the raw initial
    state between sn.reset() and before_action(0).  Returns (obs, init) like env.reset()."""
    if seed is not None:
        np.random.seed(seed)
    env.terminal = False
    env.sn.reset()
    env.period = 0
    init = read_initial_state(env)
    env.sn.before_action(0)
    names = env.sn.internal_nodes if env.global_observable else env.sn.agent_managed_facilities
    obs = np.array([list(env.sn.get_state(a).values()) for a in names], dtype=np.float32).flatten()
    return obs, init


def rollout(env, actions, seed=None, collect_costs=False):
    """Run the reference for len(actions) steps.  Returns dict(obs0, init, obs[T], rew[T], done[T])."""
    obs0, init = reset_and_capture(env, seed)
    obs_l, rew_l, done_l = [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for a in actions:
            o, r, d, _ = env.step(a)
            obs_l.append(o)
            rew_l.append(r)
            done_l.append(d)
            if d:
                break
    out = dict(obs0=obs0, init=init, obs=np.stack(obs_l), rew=np.array(rew_l, dtype=np.float64),
               done=np.array(done_l))
    if collect_costs:
        names = [n for n in NODE_NAMES if n in env.sn.nodes]
        h = np.array([env.sn.nodes[n].holding_cost_history for n in names], dtype=np.float64).T
        b = np.array([env.sn.nodes[n].backorder_cost_history for n in names], dtype=np.float64).T
        out["holding"], out["backorder"] = h, b
    return out


def make_custom_chain_env(info_lt, ship_lt, agents, global_observable=False, max_episode_steps=100,
                          holding=(0.5, 0.5, 0.5, 0.5), backorder=(1, 1, 1, 1), targets=(19, 20, 20, 14),
                          init_inventory=(0, 25), init_pipeline=(0, 9), dplusa=None):
    """A 4-echelon serial chain like envs.py:289-414 but with per-arc lead times chosen by the caller,
    built directly from the reference's Node / Arc / SupplyNetwork / InventoryManagementEnvMultiPlayer
    classes (random initial state, uniform demand 0..8).  info_lt / ship_lt are indexed by arc k whose
    customer is node k (retailer first)."""
    ref = import_reference()
    E, NC, SN, HE, DE = ref["envs"], ref["nc"], ref["sn"], ref["heuristics"], ref["demands"]
    import gym
    idx = {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1, "latest_demand": 2}
    pol = [HE.BaseStockPolicy(target_levels=[t], array_index=idx, state_dim_per_facility=7) for t in targets]
    demand = DE.Demand("uniform", low=0, high=8, size=max_episode_steps)
    nf = len(info_lt)                            # chain length: the first nf of retailer..manufacturer
    chain = NODE_NAMES[:nf]
    nodes = []
    for k, name in enumerate(chain):
        kw = dict(name=name, initial_inventory=list(init_inventory), holding_cost=holding[k], backorder_cost=backorder[k],
                  policy=pol[k])
        if k == 0:
            kw.update(is_demand_source=True, demands=demand)
        nodes.append(NC.Node(**kw))
    nodes.append(NC.Node(name="external_supplier", is_external_supplier=True))
    sup = chain[1:] + ["external_supplier"]
    arcs = [NC.Arc(sup[k], chain[k], info_lt[k], ship_lt[k], initial_shipments=list(init_pipeline),
                   initial_sales_orders=list(init_pipeline), random_init=True) for k in range(nf - 1, -1, -1)]
    names = [chain[k] for k in agents]
    net = SN.SupplyNetwork(nodes=nodes, arcs=arcs, agent_managed_facilities=names)
    n_obs = nf if global_observable else len(names)
    env = E.InventoryManagementEnvMultiPlayer(
        net, max_episode_steps=max_episode_steps, action_space=gym.spaces.Box(0, 16, shape=(len(names),)),
        observation_space=gym.spaces.Box(low=-np.inf, high=np.inf, shape=(7 * n_obs,)), global_observable=global_observable)
    if dplusa is not None:
        env = ref["wrappers"].wrap_action_d_plus_a(env, offset=dplusa[0], lb=dplusa[1], ub=dplusa[2])
    return env


# ---------------------------------------------------------------------- distribution trees
def make_network_env(nodes, arcs, agents, global_observable=False, max_episode_steps=100,
                     init_inventory=(0, 25), init_pipeline=(0, 9), dplusa=None, external_name="external_supplier"):
    """Any network the reference's classes accept, built like envs.py:289-414 builds the chain.
    nodes: list of dict(name, target, holding, backorder, demand=None | (low, high))  (internal nodes, node-list
    order; the external supplier is appended); arcs: list of (source, target, info_lt, ship_lt) in arc-list order;
    agents: node names in caller order."""
    ref = import_reference()
    E, NC, SN, HE, DE = ref["envs"], ref["nc"], ref["sn"], ref["heuristics"], ref["demands"]
    import gym
    idx = {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1, "latest_demand": 2}
    objs = []
    for nd in nodes:
        kw = dict(name=nd["name"], initial_inventory=list(init_inventory), holding_cost=nd["holding"],
                  backorder_cost=nd["backorder"],
                  policy=HE.BaseStockPolicy(target_levels=[nd["target"]], array_index=idx, state_dim_per_facility=7))
        if nd.get("demand") is not None:
            lo, hi = nd["demand"]
            kw.update(is_demand_source=True, demands=DE.Demand("uniform", low=lo, high=hi, size=max_episode_steps))
        objs.append(NC.Node(**kw))
    objs.append(NC.Node(name=external_name, is_external_supplier=True))
    arc_objs = [NC.Arc(s, t, li, ls, initial_shipments=list(init_pipeline), initial_sales_orders=list(init_pipeline),
                       random_init=True) for (s, t, li, ls) in arcs]
    net = SN.SupplyNetwork(nodes=objs, arcs=arc_objs, agent_managed_facilities=list(agents))
    n_obs = len(nodes) if global_observable else len(agents)
    env = E.InventoryManagementEnvMultiPlayer(
        net, max_episode_steps=max_episode_steps, action_space=gym.spaces.Box(0, 16, shape=(len(agents),)),
        observation_space=gym.spaces.Box(low=-np.inf, high=np.inf, shape=(7 * n_obs,)), global_observable=global_observable)
    if dplusa is not None:
        env = ref["wrappers"].wrap_action_d_plus_a(env, offset=dplusa[0], lb=dplusa[1], ub=dplusa[2])
    return env


def read_network_state(env, names):
    """(inv0[N], so[N][*], sh[N][*], demands[n_sources][T+3]) of a reference env right after sn.reset(); per node
    the pipelines of its (single) upstream arc, nearest first; demand streams in node order of the sources."""
    sn = env.sn
    inv0 = [sn.nodes[n].current_inventory for n in names]
    so, sh = [], []
    for n in names:
        (sup,) = sn.suppliers[n]
        arc = sn.arcs[(sup, n)]
        so.append([o.order_quantity for o in sorted(arc.sales_orders, key=lambda o: o.remaining_lead_time)])
        sh.append([s.quantity for s in sorted(arc.shipments, key=lambda s: s.time_till_arrival)])
    dem = [np.asarray(sn.nodes[n].demands._demands).copy() for n in names if sn.nodes[n].is_demand_source]
    return inv0, so, sh, dem


def network_rollout(env, names, actions, seed):
    """Like rollout() for a network env: dict(obs0, init, obs[T], rew[T], holding[T,N], backorder[T,N])."""
    np.random.seed(seed)
    env.terminal = False
    env.sn.reset()
    env.period = 0
    init = read_network_state(env, names)
    env.sn.before_action(0)
    obs_names = env.sn.internal_nodes if env.global_observable else env.sn.agent_managed_facilities
    obs0 = np.array([list(env.sn.get_state(a).values()) for a in obs_names], dtype=np.float32).flatten()
    obs_l, rew_l = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for a in actions:
            o, r, d, _ = env.step(a)
            obs_l.append(o)
            rew_l.append(r)
            if d:
                break
    h = np.array([env.sn.nodes[n].holding_cost_history for n in names], dtype=np.float64).T
    b = np.array([env.sn.nodes[n].backorder_cost_history for n in names], dtype=np.float64).T
    return dict(obs0=obs0, init=init, obs=np.stack(obs_l), rew=np.array(rew_l, dtype=np.float64), holding=h, backorder=b,
                order_sequence=list(env.sn.order_sequence))
