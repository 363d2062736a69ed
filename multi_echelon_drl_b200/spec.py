"""Scenario description for the batched beer-game simulator.

`ScenarioSpec` carries what the reference hard-codes in its env factories (envs.py:163-286
"normal"/basic, :289-414 "uniform"/complex, :501-596 classic) plus the
ordering rule of utils/wrappers.py.  It is lowered to the C struct `sn_spec` of
include/supplynet.h for every kernel launch.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional, Tuple

from . import _lib

NODE_NAMES = ("retailer", "wholesaler", "distributor", "manufacturer")   # order_sequence, utils/graph.py:5-32
INFO_LEADTIME = (2, 2, 2, 1)     # arc k: customer = node k (envs.py:355-392, beergame.yaml:26-50)
SHIP_LEADTIME = (2, 2, 2, 2)
OBS_PER_FACILITY = 7             # on_hand, unfilled_demand, latest_demand, unreceived_pipeline_0..3


def node_index(name_or_index) -> int:
    if isinstance(name_or_index, str):
        try:
            return NODE_NAMES.index(name_or_index)
        except ValueError:
            raise ValueError(f"unknown facility {name_or_index!r}; expected one of {NODE_NAMES}") from None
    k = int(name_or_index)
    if not 0 <= k < 4:
        raise ValueError(f"facility index {k} out of range")
    return k


@dataclass(frozen=True)
class ScenarioSpec:
    name: str
    holding_cost: Tuple[float, ...]
    backorder_cost: Tuple[float, ...]
    coplayer_target: Tuple[float, ...]
    agents: Tuple[int, ...] = (0, 1, 2, 3)          # action-column order (supplynetwork.py:217)
    global_observable: bool = False
    max_episode_steps: int = 100
    coplayer_lb: float = 0.0                         # BaseStockPolicy defaults, utils/heuristics.py:16
    coplayer_ub: float = 16.0
    rule: str = "a"                                  # 'a' | 'd+a'
    dpa_offset: float = 0.0
    dpa_lb: float = 0.0
    dpa_ub: float = 16.0
    action_low: float = 0.0                          # action space bounds (envs.py:269-272, 397-400)
    action_high: float = 16.0
    # demand / initial-state generators (host side, utils/demands.py + Arc/Node.reset)
    demand_pattern: str = "uniform"                  # 'uniform' | 'normal' | 'classic_beer_game' | 'samples' | 'list'
    demand_list: Tuple[int, ...] = ()                # 'list': the fixed sequence every episode replays (utils/demands.py:38-39)
    demand_params: Tuple[float, float] = (0, 8)      # (low, high) or (mean, sd)
    demand_path: Optional[str] = None                # .npy of demand rows for 'samples' (utils/demands.py:26-29)
    random_init: bool = True
    init_inventory: Tuple[int, int] = (0, 25)        # randint(lo, hi) if random_init else lo
    init_pipeline: Tuple[int, int] = (0, 9)          # shipments and sales orders alike
    fixed_inventory: int = 12
    fixed_shipments: Tuple[int, int] = (4, 4)
    fixed_sales_orders: Tuple[int, int] = (4, 4)
    # multi-item: cost vectors of items 1.. (item 0 uses holding_cost / backorder_cost); chains are
    # interleaved item-minor, so obs/action rows of one env are the concatenation over its items
    extra_item_costs: Tuple[Tuple[Tuple[float, ...], Tuple[float, ...]], ...] = ()
    # per-arc lead times (arc k: customer = node k).  The beer game's values select the tuned kernels,
    # anything else (1..4, info + shipment <= 5) the generic ones
    info_leadtime: Tuple[int, ...] = INFO_LEADTIME
    ship_leadtime: Tuple[int, ...] = SHIP_LEADTIME
    # chain length: the first n_facilities of retailer..manufacturer; the last one buys from the external supplier.
    # Per-facility tuples (costs, targets, lead times) then have n_facilities entries.  Shorter chains run on the
    # generic kernels with the missing upstream nodes as phantoms (supplynet_core.cuh, DevSpec::nf).
    n_facilities: int = 4
    # Distribution trees (topology='tree'): every node buys from ONE supplier and may serve several customers and / or
    # its own external demand - what the reference's SupplyNetwork does for such node / arc lists.  Node k is the k-th
    # internal node of the node list (observation and cost order); per-node tuples are indexed by it and lead times
    # belong to the node's upstream arc.  See include/supplynet.h (SN_TOPO_TREE) and csrc/supplynet_tree.cuh.
    topology: str = "chain"
    names: Tuple[str, ...] = ()                      # tree: internal node names, node-list order
    suppliers: Tuple[int, ...] = ()                  # tree: supplier node of node k, -1 = the external supplier
    demand_sources: Tuple[bool, ...] = ()            # tree: Node.is_demand_source
    arc_order: Tuple[int, ...] = ()                  # tree: upstream arcs (named by their customer node) in arc-list order
    source_demand_params: Tuple[Tuple[float, float], ...] = ()   # tree: per demand source (node order); default demand_params
    external_name: str = "external_supplier"

    def __post_init__(self):
        nf = int(self.n_facilities)
        if self.topology not in ("chain", "tree"):
            raise ValueError(f"topology must be 'chain' or 'tree', got {self.topology!r}")
        if not 1 <= nf <= (8 if self.topology == "tree" else 4):
            raise ValueError("n_facilities must be 1..4 (chains) or 1..8 (trees)")
        if self.topology == "tree":
            self._init_tree(nf)
        ag = tuple(self.node_index(a) for a in self.agents)
        object.__setattr__(self, "agents", ag)
        if not 1 <= len(ag) <= nf or len(set(ag)) != len(ag) or max(ag) >= nf:
            raise ValueError(f"agent_managed_facilities {self.agents} must name 1..{nf} distinct facilities of the network")
        pos = sorted(self.order_pos[a] for a in ag)
        if pos[-1] - pos[0] + 1 != len(ag):
            # the reference never lets the in-between node order and crashes at supplynetwork.py:199
            raise ValueError("agent-managed facilities must form a contiguous block of the chain "
                             "(of the order sequence, for trees)")
        if self.rule not in ("a", "d+a"):
            raise ValueError(f"ordering rule must be 'a' or 'd+a', got {self.rule!r}")
        object.__setattr__(self, "info_leadtime", tuple(int(x) for x in self.info_leadtime))
        object.__setattr__(self, "ship_leadtime", tuple(int(x) for x in self.ship_leadtime))
        for name in ("info_leadtime", "ship_leadtime", "holding_cost", "backorder_cost", "coplayer_target"):
            if len(getattr(self, name)) < nf:
                raise ValueError(f"{name} needs one entry per facility ({nf})")

    def _init_tree(self, nf):
        from .utils.graph import parse_order_sequence
        names = tuple(self.names)
        sup = tuple(int(x) for x in self.suppliers)
        if len(names) != nf or len(set(names)) != nf or len(sup) != nf or self.external_name in names:
            raise ValueError(f"a tree of {nf} facilities needs {nf} distinct names and one supplier index per node")
        if any(not (-1 <= x < nf) or x == k for k, x in enumerate(sup)):
            raise ValueError("suppliers[k] must be another node's index or -1 (the external supplier)")
        arcs = tuple(int(x) for x in self.arc_order) if len(self.arc_order) else tuple(range(nf))
        if sorted(arcs) != list(range(nf)):
            raise ValueError("arc_order must list every node's upstream arc once")
        src = tuple(bool(x) for x in self.demand_sources)
        if len(src) != nf or not any(src):
            raise ValueError("demand_sources needs one flag per node and at least one demand source")
        customers = {names[k]: [names[c] for c in arcs if sup[c] == k] for k in range(nf)}
        customers[self.external_name] = [names[c] for c in arcs if sup[c] < 0]
        seq = parse_order_sequence(list(names) + [self.external_name], customers)      # utils/graph.py:5-32
        internal = [n for n in seq if n != self.external_name]
        order_pos = tuple(internal.index(n) for n in names)
        if any(sup[k] >= 0 and order_pos[sup[k]] <= order_pos[k] for k in range(nf)):
            raise ValueError("the supplier relation has a cycle")
        rank = tuple([c for c in arcs if sup[c] == sup[k]].index(k) for k in range(nf))
        n_src = sum(src)
        sdp = tuple(tuple(p) for p in self.source_demand_params) or (tuple(self.demand_params),) * n_src
        if len(sdp) != n_src:
            raise ValueError("source_demand_params needs one (a, b) per demand source")
        for k, v in (("names", names), ("suppliers", sup), ("arc_order", arcs), ("demand_sources", src),
                     ("source_demand_params", sdp), ("_order_pos", order_pos), ("_customer_rank", rank)):
            object.__setattr__(self, k, v)

    # ------------------------------------------------------------------ derived
    def node_index(self, name_or_index) -> int:
        if self.topology == "tree":
            if isinstance(name_or_index, str):
                if name_or_index not in self.names:
                    raise ValueError(f"unknown facility {name_or_index!r}; expected one of {self.names}")
                return self.names.index(name_or_index)
            k = int(name_or_index)
            if not 0 <= k < self.n_facilities:
                raise ValueError(f"facility index {k} out of range")
            return k
        return node_index(name_or_index)

    @property
    def is_tree(self) -> bool:
        return self.topology == "tree"

    @property
    def order_pos(self) -> Tuple[int, ...]:
        """Position of node k among the internal nodes in the order sequence (chains: k itself)."""
        return self._order_pos if self.is_tree else tuple(range(self.n_facilities))

    @property
    def order_sequence(self) -> Tuple[str, ...]:
        pos = self.order_pos
        return tuple(self.node_names[pos.index(i)] for i in range(self.n_facilities))

    @property
    def customer_rank(self) -> Tuple[int, ...]:
        return self._customer_rank if self.is_tree else (0,) * self.n_facilities

    @property
    def source_nodes(self) -> Tuple[int, ...]:
        """Nodes with an external demand stream, in node order (= row order of the demand tables)."""
        return tuple(k for k in range(self.n_facilities) if self.demand_sources[k]) if self.is_tree else (0,)

    @property
    def n_sources(self) -> int:
        return len(self.source_nodes)

    @property
    def standard_leadtimes(self) -> bool:
        """True for the beer game itself (4 facilities, lead times 2,2,2,1 / 2,2,2,2): tuned kernels, packed tables."""
        return (not self.is_tree and self.n_facilities == 4 and self.info_leadtime == INFO_LEADTIME
                and self.ship_leadtime == SHIP_LEADTIME)

    @property
    def node_names(self) -> Tuple[str, ...]:
        return tuple(self.names) if self.is_tree else NODE_NAMES[: self.n_facilities]

    @property
    def n_items(self) -> int:
        return 1 + len(self.extra_item_costs)

    @property
    def n_agents(self) -> int:
        return len(self.agents)

    @property
    def obs_nodes(self) -> Tuple[int, ...]:
        return tuple(range(self.n_facilities)) if self.global_observable else self.agents

    @property
    def obs_dim(self) -> int:
        return OBS_PER_FACILITY * len(self.obs_nodes)

    def with_rule(self, rule: str, offset: float, lb: float, ub: float) -> "ScenarioSpec":
        return replace(self, rule=rule, dpa_offset=offset, dpa_lb=lb, dpa_ub=ub)

    def to_c(self) -> _lib.SnSpec:
        s = _lib.SnSpec()
        s.n_facilities = self.n_facilities
        for k in range(self.n_facilities):
            s.info_leadtime[k] = self.info_leadtime[k]
            s.ship_leadtime[k] = self.ship_leadtime[k]
            s.holding_cost[k] = float(self.holding_cost[k])
            s.backorder_cost[k] = float(self.backorder_cost[k])
            s.coplayer_target[k] = float(self.coplayer_target[k])
        for k in range(_lib.MAXF):
            s.agents[k] = self.agents[k] if k < len(self.agents) else 0
        s.n_agents = len(self.agents)
        s.global_observable = int(self.global_observable)
        s.max_episode_steps = int(self.max_episode_steps)
        s.rule = _lib.SN_RULE_D_PLUS_A if self.rule == "d+a" else _lib.SN_RULE_A
        s.coplayer_lb, s.coplayer_ub = float(self.coplayer_lb), float(self.coplayer_ub)
        s.dpa_offset, s.dpa_lb, s.dpa_ub = float(self.dpa_offset), float(self.dpa_lb), float(self.dpa_ub)
        if len(self.extra_item_costs) > 3:
            raise ValueError("at most 4 items")
        s.n_items = self.n_items
        if self.is_tree:
            s.topology = _lib.SN_TOPO_TREE
            for k in range(self.n_facilities):
                s.supplier[k] = self.suppliers[k]
                s.demand_source[k] = int(self.demand_sources[k])
                s.order_pos[k] = self.order_pos[k]
                s.customer_rank[k] = self.customer_rank[k]
        for i, (h, b) in enumerate(self.extra_item_costs):
            for k in range(self.n_facilities):
                s.item_holding_cost[i][k] = float(h[k])
                s.item_backorder_cost[i][k] = float(b[k])
        return s


def _agents(agent_managed_facilities, default):
    if agent_managed_facilities is None:
        agent_managed_facilities = default
    return tuple(node_index(a) for a in agent_managed_facilities)


def uniform_spec(agent_managed_facilities=None, global_observable=False, max_episode_steps=100,
                 random_init=True) -> ScenarioSpec:
    """The "complex" scenario, make_beer_game_uniform_multi_facility (envs.py:289-414)."""
    return ScenarioSpec(
        name="uniform", holding_cost=(0.5,) * 4, backorder_cost=(1.0,) * 4, coplayer_target=(19, 20, 20, 14),
        agents=_agents(agent_managed_facilities, NODE_NAMES), global_observable=global_observable,
        max_episode_steps=max_episode_steps, action_low=0, action_high=16, demand_pattern="uniform",
        demand_params=(0, 8), random_init=random_init, init_inventory=(0, 25), init_pipeline=(0, 9),
        fixed_inventory=12, fixed_shipments=(4, 4), fixed_sales_orders=(4, 4))


def normal_spec(agent_managed_facilities=None, global_observable=False, max_episode_steps=100,
                random_init=False) -> ScenarioSpec:
    """The "basic" scenario, make_beer_game_normal_multi_facility (envs.py:163-286).  Co-players
    keep the default clip [0, 16] although the action range is [0, 20]."""
    return ScenarioSpec(
        name="normal", holding_cost=(1.0, 0.75, 0.5, 0.25), backorder_cost=(10.0, 0.0, 0.0, 0.0),
        coplayer_target=(48, 43, 41, 30), agents=_agents(agent_managed_facilities, ("retailer",)),
        global_observable=global_observable, max_episode_steps=max_episode_steps, action_low=0, action_high=20,
        demand_pattern="normal", demand_params=(10, 2), random_init=random_init, init_inventory=(0, 21),
        init_pipeline=(0, 21), fixed_inventory=10, fixed_shipments=(10, 0), fixed_sales_orders=(10, 0))


def classic_spec(agent_managed_facilities=None, max_episode_steps=35, demand_pattern="classic_beer_game",
                 demand_list=()) -> ScenarioSpec:
    """make_beer_game (envs.py:501-596): fixed initial state only (random_init is never wired, Q16)."""
    params = {"classic_beer_game": (4, 8), "normal": (10, 2), "uniform": (0, 2), "list": (0, 0)}[demand_pattern]
    return ScenarioSpec(
        name="classic", holding_cost=(0.5,) * 4, backorder_cost=(1.0,) * 4, coplayer_target=(32, 32, 32, 24),
        agents=_agents(agent_managed_facilities, ("retailer",)), global_observable=False,
        max_episode_steps=max_episode_steps, action_low=0, action_high=16, demand_pattern=demand_pattern,
        demand_params=params, demand_list=tuple(int(x) for x in demand_list), random_init=False, fixed_inventory=12,
        fixed_shipments=(4, 4), fixed_sales_orders=(4, 4))


def tree_spec(nodes, arcs, agent_managed_facilities, global_observable=False, max_episode_steps=100, random_init=True,
              demand_pattern="uniform", name="tree", **overrides) -> ScenarioSpec:
    """Distribution-tree scenario from node / arc lists shaped like the reference's constructors take them
    (SupplyNetwork(nodes=[Node...], arcs=[Arc...]), envs.py:289-414):
      nodes  list of dict(name, holding_cost, backorder_cost, target [co-player base-stock level],
             demand=None | (a, b) [external demand stream: uniform low/high or normal mean/sd]) - internal nodes in
             node-list order;
      arcs   list of (supplier_name, customer_name, info_leadtime, shipment_leadtime) in arc-list order; a supplier
             name that is not a node is the external supplier.
    Up to eight internal nodes (more than four run on the shared-memory kernels), one supplier per node."""
    names = tuple(n["name"] for n in nodes)
    nf = len(names)
    if not 1 <= nf <= 8:
        raise NotImplementedError("kernels hold up to eight internal nodes per network")
    up = {}
    for (s_name, c_name, li, ls) in arcs:
        if c_name not in names:
            raise ValueError(f"arc customer {c_name!r} is not an internal node")
        if c_name in up:
            raise NotImplementedError(f"node {c_name!r} has more than one supplier; only distribution trees are supported")
        up[c_name] = (names.index(s_name) if s_name in names else -1, int(li), int(ls))
    if set(up) != set(names):
        raise ValueError("every internal node needs exactly one upstream arc")
    ext = [s_name for (s_name, _, _, _) in arcs if s_name not in names]
    u = uniform_spec(None, global_observable, max_episode_steps, random_init)
    src = tuple(n.get("demand") is not None for n in nodes)
    spec = replace(
        u, name=name, topology="tree", n_facilities=nf, names=names, suppliers=tuple(up[n][0] for n in names),
        info_leadtime=tuple(up[n][1] for n in names), ship_leadtime=tuple(up[n][2] for n in names),
        arc_order=tuple(names.index(c) for (_, c, _, _) in arcs), demand_sources=src, demand_pattern=demand_pattern,
        source_demand_params=tuple(tuple(n["demand"]) for n in nodes if n.get("demand") is not None),
        holding_cost=tuple(float(n.get("holding_cost", 0.5)) for n in nodes),
        backorder_cost=tuple(float(n.get("backorder_cost", 1.0)) for n in nodes),
        coplayer_target=tuple(float(n.get("target", 20)) for n in nodes),
        external_name=ext[0] if ext else "external_supplier",
        agents=tuple(agent_managed_facilities if agent_managed_facilities is not None else range(nf)), **overrides)
    return spec


def _tree_from_config(cfg, agent_managed_facilities, global_observable, max_episode_steps, coplayer_target, random_init):
    """network_configs/*.yaml schema -> tree_spec.  A node with `demand_path` is a demand source; here the value
    names a generator 'uniform:low:high' / 'normal:mean:sd' (the reference's sample files are not in its repo);
    `target` per node (or coplayer_target by position) is the co-player base-stock level."""
    if len(cfg.get("items", []) or []) > 1:
        raise NotImplementedError("multi-item networks are supported for the serial beer-game chain only")
    internal = [n for n in cfg["nodes"] if not n.get("is_external_supplier", False)]
    pattern, nodes = None, []
    for i, n in enumerate(internal):
        demand = None
        if n.get("demand_path"):
            parts = str(n["demand_path"]).split(":")
            if len(parts) != 3 or parts[0] not in ("uniform", "normal"):
                raise NotImplementedError(f"demand_path {n['demand_path']!r}: expected 'uniform:low:high' or 'normal:mean:sd'")
            if pattern not in (None, parts[0]):
                raise NotImplementedError("all demand sources of a network must use the same demand pattern")
            pattern, demand = parts[0], (float(parts[1]), float(parts[2]))
        cost = lambda v: float(v[0] if isinstance(v, (list, tuple)) else v)
        nodes.append(dict(name=n["name"], holding_cost=cost(n.get("holding_cost", 0)), backorder_cost=cost(n.get("backorder_cost", 0)),
                          target=n.get("target", coplayer_target[min(i, len(coplayer_target) - 1)]), demand=demand))
    arcs = [(a["supplier"], a["customer"], int(a.get("info_leadtime", 0)), int(a.get("shipment_leadtime", 0))) for a in cfg["arcs"]]
    if any(not 1 <= l <= 4 for a in arcs for l in a[2:]) or any(a[2] + a[3] > 5 for a in arcs):
        raise NotImplementedError("lead times must be 1..4 with info + shipment <= 5 per arc")
    return tree_spec(nodes, arcs, agent_managed_facilities, global_observable, max_episode_steps, random_init,
                     demand_pattern=pattern or "uniform", name="config-tree")


def specs_from_config(network_config: dict, agent_managed_facilities=None, global_observable=False,
                      max_episode_steps=100, coplayer_target=(19, 20, 20, 14), random_init=True):
    """ScenarioSpec per item from a network description (network_configs/*.yaml schema; what the
    reference's `supplynetwork.from_dict` was meant to do, supplynetwork.py:345-397).  Multi-item
    descriptions (`items:`) give one spec per item: items are independent copies of the chain with
    their own cost vectors (the reference defines no item coupling)."""
    from .utils.graph import chain_from_config
    try:
        seq, info, ship, hold, back = chain_from_config(network_config)
        if tuple(seq) != NODE_NAMES[:len(seq)]:
            raise NotImplementedError("not the beer game's node names")
    except NotImplementedError:
        # not the serial chain retailer..manufacturer: a distribution tree (any names), if it is one
        return [_tree_from_config(network_config, agent_managed_facilities, global_observable, max_episode_steps,
                                  coplayer_target, random_init)]
    nf = len(seq)
    if not 1 <= nf <= 4:
        raise NotImplementedError(f"kernels implement serial chains made of the first 1..4 of {NODE_NAMES}; got {seq}")
    if any(not 1 <= l <= 4 for l in info + ship) or any(a + b > 5 for a, b in zip(info, ship)):
        raise NotImplementedError(f"lead times must be 1..4 with info + shipment <= 5 per arc; got {info}/{ship}")
    agents = _agents(agent_managed_facilities, NODE_NAMES[:nf])
    u = uniform_spec(None, global_observable, max_episode_steps, random_init)
    return [replace(u, name=f"config-item{i}", n_facilities=nf, agents=agents, info_leadtime=info, ship_leadtime=ship,
                    holding_cost=tuple(float(x) for x in hold[i]), backorder_cost=tuple(float(x) for x in back[i]),
                    coplayer_target=tuple(coplayer_target)[:nf])
            for i in range(len(hold))]
