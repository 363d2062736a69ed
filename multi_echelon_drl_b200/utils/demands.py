"""Host-side demand and initial-state streams (mirror of the reference's utils/demands.py).

Per north_star the random draws stay on the host so that inputs match the reference: the
reference draws everything from numpy's legacy global RNG (utils/demands.py:44,47;
network_components.py:143,157,297).  Two generators are offered:

  * `seeded_episode_inputs(spec, seeds)`  - one `np.random.RandomState(seed)` per env, consumed in
    exactly the order of `np.random.seed(seed); env.reset()` in the reference, so env i
    reproduces the reference env seeded with seeds[i].  ~150 us per env: for parity subsets.
  * `bulk_episode_inputs(spec, n, rng)`   - the same distributions drawn with a few vectorised
    calls from one stream, for throughput runs (not stream-compatible with per-env seeding).

Both return `init [n, 19]` (inv[4]; sales orders so[k][j], k=0..3, 2+2+2+1; shipments sh[k][j],
4x2 - the row order of `sn_reset`) and `demand [n, T+3]` (utils/demands.py:21 "size + 3").
"""
from __future__ import annotations

import numpy as np

from ..spec import ScenarioSpec

INIT_WORDS = 19
_SO_OFF = (4, 6, 8, 10)
_SH_OFF = (11, 13, 15, 17)


class Demand:
    """Batched counterpart of the reference's `Demand` (utils/demands.py:5-67): same constructor
    arguments and patterns; `sample(n, rng)` draws the per-episode demand arrays of n envs."""

    def __init__(self, demand_pattern="classic_beer_game", data_path=None, size=100, low=None, high=None,
                 mean=None, sd=None):
        self.demands_pattern = demand_pattern
        self.size = size + 3
        self.low, self.high, self.mean, self.sd = low, high, mean, sd
        self.samples = np.load(data_path) if data_path is not None else None
        self.sample_pointer = 0
        if isinstance(demand_pattern, str):
            if demand_pattern == "uniform" and (low is None or high is None):
                raise ValueError('"low" and "high" need to be provided when uniform distribution pattern is specified')
            if demand_pattern == "normal" and (mean is None or sd is None):
                raise ValueError('"mean" and "sd" need to be provided when normal distribution pattern is specified')
            if demand_pattern not in ("uniform", "normal", "classic_beer_game", "samples"):
                raise ValueError("Demand pattern not recognized")

    def draw(self, rs) -> np.ndarray:
        """One episode from `rs` (a RandomState or the np.random module), consuming it exactly like
        Demand.reset (utils/demands.py:37-59)."""
        p = self.demands_pattern
        if isinstance(p, (np.ndarray, list, tuple)):
            return np.asarray(p)
        if p == "uniform":
            return rs.randint(self.low, self.high + 1, self.size)
        if p == "normal":
            return np.round(np.maximum(self.mean + self.sd * rs.randn(self.size), 0)).astype(int)
        if p == "classic_beer_game":
            return np.array([4] * 4 + [8] * (self.size - 4))
        row = self.samples[self.sample_pointer % self.samples.shape[0]]      # "samples"
        self.sample_pointer += 1
        return np.asarray(row)

    def row(self, rs) -> np.ndarray:
        """`draw` as a row of exactly `size` values for the batched demand table: a list pattern (passed through as
        is by `draw`, like the reference) is cut or zero-padded - the single env raises the reference's IndexError
        before a period past the end of the list is read (gym_env.py)."""
        row = self.draw(rs)
        if row.shape[0] >= self.size:
            return row[: self.size]
        return np.concatenate([row, np.zeros(self.size - row.shape[0], dtype=row.dtype)])

    def sample(self, n: int, rs) -> np.ndarray:
        """[n, size] demands in a few vectorised draws."""
        p = self.demands_pattern
        if isinstance(p, (np.ndarray, list, tuple)):
            return np.broadcast_to(np.asarray(p), (n, len(p))).copy()
        if p == "uniform":
            return rs.randint(self.low, self.high + 1, (n, self.size))
        if p == "normal":
            return np.round(np.maximum(self.mean + self.sd * rs.randn(n, self.size), 0)).astype(int)
        if p == "classic_beer_game":
            return np.broadcast_to(np.array([4] * 4 + [8] * (self.size - 4)), (n, self.size)).copy()
        rows = [self.draw(rs)[: self.size] for _ in range(n)]
        return np.stack(rows)


_SAMPLE_DEMANDS = {}


def demand_for(spec: ScenarioSpec, source: int = 0) -> Demand:
    """Demand object of the `source`-th demand source (trees may give every source its own parameters)."""
    if spec.is_tree and spec.demand_pattern in ("uniform", "normal"):
        a, b = spec.source_demand_params[source]
        return (Demand("uniform", low=int(a), high=int(b), size=spec.max_episode_steps) if spec.demand_pattern == "uniform"
                else Demand("normal", mean=a, sd=b, size=spec.max_episode_steps))
    if spec.demand_pattern == "samples":       # one object per file: its row pointer advances across resets
        key = spec.demand_path
        if key not in _SAMPLE_DEMANDS:
            _SAMPLE_DEMANDS[key] = Demand("samples", data_path=spec.demand_path, size=spec.max_episode_steps)
        return _SAMPLE_DEMANDS[key]
    a, b = spec.demand_params
    if spec.demand_pattern == "list":
        return Demand(list(spec.demand_list), size=spec.max_episode_steps)
    if spec.demand_pattern == "uniform":
        return Demand("uniform", low=int(a), high=int(b), size=spec.max_episode_steps)
    if spec.demand_pattern == "normal":
        return Demand("normal", mean=a, sd=b, size=spec.max_episode_steps)
    if spec.demand_pattern == "samples":
        return Demand("samples", data_path=spec.demand_path, size=spec.max_episode_steps)
    return Demand(spec.demand_pattern, size=spec.max_episode_steps)


def init_layout(spec: ScenarioSpec):
    """(n_words, so_offsets[4], sh_offsets[4]) of the init rows: packed 19 words for the beer game's lead
    times (tuned kernels), 36 words (4 slots per pipeline) for any other lead times (generic kernels)."""
    if spec.standard_leadtimes:
        return INIT_WORDS, _SO_OFF, _SH_OFF
    if spec.n_facilities > 4:                    # trees of 5..8 nodes: inv[8]; so[k][0..3] at 8+4k; sh[k][0..3] at 40+4k
        return 72, tuple(8 + 4 * k for k in range(8)), tuple(40 + 4 * k for k in range(8))
    return 36, tuple(4 + 4 * k for k in range(4)), tuple(20 + 4 * k for k in range(4))


def _fixed_init(spec: ScenarioSpec) -> np.ndarray:
    n_words, so_off, sh_off = init_layout(spec)
    v = np.zeros(n_words, dtype=np.int64)
    v[:spec.n_facilities] = spec.fixed_inventory
    for k in range(spec.n_facilities):
        for j in range(spec.info_leadtime[k]):
            v[so_off[k] + j] = spec.fixed_sales_orders[min(j, len(spec.fixed_sales_orders) - 1)]
        for j in range(spec.ship_leadtime[k]):
            v[sh_off[k] + j] = spec.fixed_shipments[min(j, len(spec.fixed_shipments) - 1)]
    return v


def seeded_episode_inputs(spec: ScenarioSpec, seeds):
    """Per-env streams in the reference's draw order: inv_retailer, demand[T+3], inv_wholesaler,
    inv_distributor, inv_manufacturer, then per arc in the factories' arc-list order
    (ext->man, man->dis, dis->who, who->ret): shipments, then sales orders."""
    seeds = np.asarray(seeds).reshape(-1)
    return seeded_episode_inputs_from_streams(spec, [np.random.RandomState(int(s)) for s in seeds])


def bulk_episode_inputs(spec: ScenarioSpec, n: int, rs):
    """Same distributions, vectorised draws from one stream `rs` (RandomState).  Trees: demand is
    [n, n_sources, T+3], one stream per demand source in node order."""
    dem = demand_for(spec)
    n_words, so_off, sh_off = init_layout(spec)
    if spec.random_init:
        init = np.zeros((n, n_words), dtype=np.int64)
        n_inv = so_off[0]                           # inventory rows: 4, or 8 for trees of more than four nodes
        init[:, :n_inv] = rs.randint(spec.init_inventory[0], spec.init_inventory[1], (n, n_inv))
        init[:, spec.n_facilities:n_inv] = 0
        pipe = rs.randint(spec.init_pipeline[0], spec.init_pipeline[1], (n, n_words - n_inv))
        if not spec.standard_leadtimes:             # slots beyond an arc's lead time (or chain) stay empty
            mask = np.zeros(n_words - n_inv, dtype=bool)
            for k in range(spec.n_facilities):
                mask[so_off[k] - n_inv: so_off[k] - n_inv + spec.info_leadtime[k]] = True
                mask[sh_off[k] - n_inv: sh_off[k] - n_inv + spec.ship_leadtime[k]] = True
            pipe = pipe * mask
        init[:, n_inv:] = pipe
    else:
        init = np.broadcast_to(_fixed_init(spec), (n, n_words)).copy()
    if spec.is_tree:
        return init, np.stack([demand_for(spec, s).sample(n, rs) for s in range(spec.n_sources)], axis=1)
    return init, dem.sample(n, rs)


def seeded_episode_inputs_from_streams(spec: ScenarioSpec, streams):
    """Like seeded_episode_inputs but continuing persistent per-env RandomState objects, so that env i
    over successive resets equals one reference env seeded once and reset repeatedly."""
    n = len(streams)
    dem = demand_for(spec)
    n_words, so_off, sh_off = init_layout(spec)
    init = np.zeros((n, n_words), dtype=np.int64)
    fixed = _fixed_init(spec)
    if spec.is_tree:
        # SupplyNetwork.reset (supplynetwork.py:94-105): nodes in node-list order (inventory, then the demand
        # episode of a demand source), then arcs in arc-list order (shipments, then sales orders)
        dems = [demand_for(spec, s) for s in range(spec.n_sources)]
        demand = np.zeros((n, spec.n_sources, dem.size), dtype=np.int64)
        for i, rs in enumerate(streams):
            if not spec.random_init:
                init[i] = fixed
            for k in range(spec.n_facilities):
                if spec.random_init:
                    init[i, k] = rs.randint(*spec.init_inventory)
                if spec.demand_sources[k]:
                    s = spec.source_nodes.index(k)
                    demand[i, s] = dems[s].draw(rs)
            if spec.random_init:
                for k in spec.arc_order:
                    for j in range(spec.ship_leadtime[k]):
                        init[i, sh_off[k] + j] = rs.randint(*spec.init_pipeline)
                    for j in range(spec.info_leadtime[k]):
                        init[i, so_off[k] + j] = rs.randint(*spec.init_pipeline)
        return init, demand
    demand = np.zeros((n, dem.size), dtype=np.int64)
    for i, rs in enumerate(streams):
        if not spec.random_init:
            init[i] = fixed
            demand[i] = dem.row(rs)
            continue
        lo, hi = spec.init_inventory
        plo, phi = spec.init_pipeline
        init[i, 0] = rs.randint(lo, hi)
        demand[i] = dem.row(rs)
        for k in range(1, spec.n_facilities):
            init[i, k] = rs.randint(lo, hi)
        for k in range(spec.n_facilities - 1, -1, -1):
            for j in range(spec.ship_leadtime[k]):
                init[i, sh_off[k] + j] = rs.randint(plo, phi)
            for j in range(spec.info_leadtime[k]):
                init[i, so_off[k] + j] = rs.randint(plo, phi)
    return init, demand
