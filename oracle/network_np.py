"""CPU oracle for DISTRIBUTION TREES: numpy restatement of the reference's SupplyNetwork step for networks in
which every internal node buys from exactly one supplier (another internal node or the external supplier) and
may serve several customers and/or its own external demand (TEST INFRASTRUCTURE).

THIS IS THE CHECKER, NOT THE PRODUCT.  Only tests/ may import it; the product path never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks it against tests/golden/trees.npz (trajectories
recorded from the reference's own Node / Arc / SupplyNetwork / InventoryManagementEnvMultiPlayer classes on
tree networks, tests/golden/make_golden.py:gen_trees), against the serial-chain goldens (a chain is a tree), and
- when /root/reference is mounted - differentially against the live reference.

It follows the reference's own sequencing (all citations into /root/reference), vectorised over envs:
  order sequence   utils/graph.py:5-32 (name-sorted depth-first search, customers before suppliers);
                   customers_dict in arc-list order (supplynetwork.py:22-25)
  reset            supplynetwork.py:94-105, network_components.py:138-168,293-319, then before_action(0)
  before_action    supplynetwork.py:182-210: per node of the order sequence, on its upstream arc: advance the
                   slips, supplier.latest_demand = [arrived] (OVERWRITTEN per arc, :193 - with several customers
                   the one latest in the order sequence wins), unreceived list pop/merge; then the co-players
                   ahead of the first agent order
  agent_action     supplynetwork.py:215-227;  after_action :229-268: co-players behind the last agent order;
                   in shipment sequence (reverse order sequence) every node first lets its customers' arcs
                   receive (inventory +=, FIFO consumption of the unreceived list), then serves its customer
                   arcs IN customers_dict ORDER from what it has (Arc.fill_orders, network_components.py:203-234:
                   the first arc has priority), then its own external demand (:323-345)
  get_state        supplynetwork.py:114-171: unfilled_demand = sum over customer arcs of arrived unshipped orders
                   + current external demand + unfilled independent demand; latest_demand = sum(node.latest_demand)
                   + current external demand; first four slots of the upstream arc's unreceived list
  get_cost         supplynetwork.py:285-310 (node-list order; holding and backorder summed apart)
Multi-supplier (assembly) networks are NOT covered: the reference's state dict keys collide there (:165-167).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

EXTERNAL = -1


def order_sequence(names: Sequence[str], customers: dict) -> List[str]:
    """utils/graph.py:5-32: repeatedly start a depth-first search at the alphabetically first node that has
    not ordered yet; a node orders after all its customers (visited in customers_dict order)."""
    pending = sorted(names)
    acted = {n: False for n in pending}
    seq: List[str] = []

    def visit(n):
        for c in customers[n]:
            if not acted[c]:
                visit(c)
        acted[n] = True
        pending.remove(n)
        seq.append(n)

    while pending:
        visit(pending[0])
    return seq


@dataclass
class TreeSpec:
    names: Sequence[str]                      # internal nodes in node-list order (= observation / cost order)
    supplier: Sequence[int]                   # per node: index of its supplier, EXTERNAL for the external supplier
    demand_source: Sequence[bool]             # per node: has its own external demand stream
    info_lt: Sequence[int]                    # per node: information lead time of its upstream arc
    ship_lt: Sequence[int]
    holding_cost: Sequence[float]
    backorder_cost: Sequence[float]
    coplayer_target: Sequence[float]
    arc_order: Sequence[int] = ()             # upstream arcs (named by their customer node) in arc-list order
    agents: Sequence[int] = (0,)              # node indices, caller order
    global_observable: bool = False
    max_episode_steps: int = 100
    coplayer_lb: float = 0.0
    coplayer_ub: float = 16.0
    rule: str = "a"
    dpa_offset: float = -8.0
    dpa_lb: float = 0.0
    dpa_ub: float = 16.0
    external_name: str = "external_supplier"
    # derived
    order_pos: List[int] = field(default_factory=list)       # position of node k among internal nodes in the order sequence
    customers: List[List[int]] = field(default_factory=list)  # per node: customer nodes in customers_dict order

    def __post_init__(self):
        n = len(self.names)
        arcs = list(self.arc_order) if len(self.arc_order) else list(range(n))
        assert sorted(arcs) == list(range(n))
        self.arc_order = arcs
        self.customers = [[c for c in arcs if self.supplier[c] == k] for k in range(n)]
        ext_customers = [c for c in arcs if self.supplier[c] == EXTERNAL]
        cd = {self.names[k]: [self.names[c] for c in self.customers[k]] for k in range(n)}
        cd[self.external_name] = [self.names[c] for c in ext_customers]
        seq = order_sequence(list(self.names) + [self.external_name], cd)
        internal = [s for s in seq if s != self.external_name]
        self.order_pos = [internal.index(nm) for nm in self.names]
        self.sequence = [self.names.index(nm) for nm in internal]          # node indices in order sequence
        pos = sorted(self.order_pos[a] for a in self.agents)
        if pos != list(range(pos[0], pos[-1] + 1)):
            # supplynetwork.py:201,231: a node between the first and the last agent never orders
            raise ValueError("agent-managed facilities must be contiguous in the order sequence")

    def obs_nodes(self):
        return list(range(len(self.names))) if self.global_observable else list(self.agents)

    @property
    def sources(self):
        return [k for k in range(len(self.names)) if self.demand_source[k]]


class TreeOracle:
    def __init__(self, spec: TreeSpec, n_envs: int, dtype=np.int64):
        self.spec, self.n, self.dtype = spec, n_envs, np.dtype(dtype)
        self.N = len(spec.names)
        self.t, self.terminal = 0, False

    # ------------------------------------------------------------------ reset
    def reset(self, inv0, so, sh, demand):
        """inv0 [n, N]; so / sh: per node k arrays [n, li_k] / [n, ls_k] of its upstream arc (nearest first);
        demand [n, n_sources, T+3] in the order of spec.sources."""
        s, dt, n = self.spec, self.dtype, self.n
        z = lambda: np.zeros(n, dtype=dt)
        self.inv = [np.array(inv0[:, k], dtype=dt) for k in range(self.N)]
        self.info = [[np.array(so[k][:, j], dtype=dt) for j in range(s.info_lt[k])] for k in range(self.N)]
        self.ship = [[np.array(sh[k][:, j], dtype=dt) for j in range(s.ship_lt[k])] for k in range(self.N)]
        self.U = []
        for k in range(self.N):                       # network_components.py:168
            full = [z() for _ in range(4)] + [a.copy() for a in self.ship[k]] + [a.copy() for a in self.info[k]]
            self.U.append(full[::-1][:5])
        self.B = [z() for _ in range(self.N)]         # arrived, unshipped orders on node k's upstream arc
        self.latest = [z() for _ in range(self.N)]    # sum(node.latest_demand): [] at reset
        self.unf_ind = [z() for _ in range(self.N)]
        self.unf_post = [z() for _ in range(self.N)]
        self.demand = np.asarray(demand)
        self.cur = [z() for _ in range(self.N)]
        for i, k in enumerate(s.sources):
            self.cur[k] = self.demand[:, i, 0].astype(dt)
        self.t, self.terminal = 0, False
        self._before()
        return self.obs()

    # ------------------------------------------------------------------ pieces
    def _before(self):
        s = self.spec
        for k in s.sequence:                          # supplynetwork.py:185-199
            a = self.info[k].pop(0)
            self.info[k].append(np.zeros(self.n, dtype=self.dtype))
            if s.supplier[k] != EXTERNAL:
                self.latest[s.supplier[k]] = a        # overwritten per arc (:193)
            self.B[k] = self.B[k] + a
            last = self.U[k].pop()
            self.U[k][-1] = self.U[k][-1] + last

    def _unfilled(self, k):
        tot = self.cur[k] + self.unf_ind[k]
        for c in self.spec.customers[k]:
            tot = tot + self.B[c]
        return tot

    def _latest(self, k):
        return self.latest[k] + self.cur[k]

    def _coplayer_order(self, k):
        s = self.spec
        on_order = self.U[k][0] + self.U[k][1] + self.U[k][2] + self.U[k][3]
        q = s.coplayer_target[k] - (self.inv[k] + on_order - self._unfilled(k))
        return np.clip(q, s.coplayer_lb, s.coplayer_ub).astype(self.dtype)

    def _place(self, k, q):
        q = np.maximum(q, 0).astype(self.dtype)
        self.info[k][self.spec.info_lt[k] - 1] = q
        self.U[k].insert(0, q)

    def _node_state(self, k, placed_pre):
        if placed_pre and len(self.U[k]) == 4:        # ordered inside before_action: 5-slot list, first four read
            u = [np.maximum(self._coplayer_order(k), 0)] + self.U[k][:3]
        else:
            u = self.U[k][:4]
        return [self.inv[k], self._unfilled(k), self._latest(k)] + list(u)

    def obs(self):
        s = self.spec
        first = min(s.order_pos[a] for a in s.agents)
        cols = []
        for k in s.obs_nodes():
            cols += self._node_state(k, placed_pre=(s.order_pos[k] < first) and not self.terminal)
        return np.stack(cols, axis=1).astype(np.float32)

    # ------------------------------------------------------------------ step
    def step(self, actions):
        if self.terminal:
            raise ValueError("Cannot take action when the state is terminal.")
        s = self.spec
        actions = np.asarray(actions)
        ag = list(s.agents)
        if s.rule == "d+a":
            lat = np.stack([self._latest(k) for k in ag], axis=1)
            actions = np.clip(lat + actions + s.dpa_offset, s.dpa_lb, s.dpa_ub)
        orders = {k: (actions[:, ag.index(k)] if k in ag else self._coplayer_order(k)) for k in range(self.N)}
        for k in range(self.N):
            self._place(k, orders[k])

        for node in [EXTERNAL] + s.sequence[::-1]:    # shipment sequence; the external supplier's place in it does
            custs = ([c for c in s.arc_order if s.supplier[c] == EXTERNAL] if node == EXTERNAL else s.customers[node])
            for c in custs:                           # not matter: it only serves arcs, from unlimited stock
                arr = self.ship[c].pop(0)
                self.ship[c].append(np.zeros(self.n, dtype=self.dtype))
                self.inv[c] = self.inv[c] + arr
                rem = arr.copy()
                for i in range(len(self.U[c]) - 1, -1, -1):
                    f = np.minimum(self.U[c][i], rem)
                    self.U[c][i] = self.U[c][i] - f
                    rem = rem - f
            unf = np.zeros(self.n, dtype=self.dtype)
            for c in custs:                           # customers_dict order: the first arc is served first
                if node == EXTERNAL:
                    q = self.B[c].copy()
                else:
                    q = np.minimum(self.inv[node], self.B[c])
                    self.inv[node] = self.inv[node] - q
                self.ship[c][s.ship_lt[c] - 1] = self.ship[c][s.ship_lt[c] - 1] + q
                self.B[c] = self.B[c] - q
                unf = unf + self.B[c]
            if node != EXTERNAL:
                if s.demand_source[node]:
                    tot = self.cur[node] + self.unf_ind[node]
                    q = np.maximum(0, np.minimum(self.inv[node], tot))
                    self.inv[node] = self.inv[node] - q
                    self.unf_ind[node] = tot - q
                    unf = unf + self.unf_ind[node]
                self.unf_post[node] = unf
        for i, k in enumerate(s.sources):
            self.cur[k] = self.demand[:, i, self.t + 1].astype(self.dtype)

        ch = np.zeros(self.n, dtype=np.float64)
        cb = np.zeros(self.n, dtype=np.float64)
        self.holding = np.zeros((self.n, self.N), dtype=np.float64)
        self.backorder = np.zeros((self.n, self.N), dtype=np.float64)
        for k in range(self.N):
            self.holding[:, k] = -self.inv[k].astype(np.float64) * s.holding_cost[k]
            self.backorder[:, k] = -self.unf_post[k].astype(np.float64) * s.backorder_cost[k]
            ch = ch + self.holding[:, k]
            cb = cb + self.backorder[:, k]
        reward = ch + cb
        self.t += 1
        if self.t >= s.max_episode_steps:
            self.terminal = True
        else:
            self._before()
        return self.obs(), reward, self.terminal

    def check_invariant(self):
        for k in range(self.N):
            lhs = sum(self.U[k])
            rhs = sum(self.info[k]) + self.B[k] + sum(self.ship[k])
            assert np.all(np.abs(lhs - rhs) < 1e-3), k
