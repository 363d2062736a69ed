"""Single-env gym-0.21 duck type on top of the CUDA simulator.

`BeerGameEnv` mirrors `InventoryManagementEnvMultiPlayer` (reference envs.py:29-117): `reset() ->
np.float32[7F']`, `step(quantity) -> (obs, float, bool, {})` (old 4-tuple API), `seed`, `action_space`,
`observation_space`, `period`, `terminal`, `max_episode_steps`, and an `sn` attribute exposing
`get_state(facility)` / `agent_managed_facilities` (what utils/wrappers.py:23-24 reads).  It is a batch of
one env on the GPU: every call launches the same kernels as the VecEnv.  Meant for drop-in use of
single-env code (the reference's own tests, SB3's Monitor/DummyVecEnv/check_env); throughput comes from
`SupplyNetVecEnv`.

Errors like the reference: ValueError when stepping a terminal env (envs.py:88-89), TypeError for
quantities that are not numbers (network_components.py:21-29).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import spaces
from .simulator import BatchedSupplyNetwork
from .spec import ScenarioSpec
from .utils import demands as _dem

_KEYS = ("on_hand", "unfilled_demand", "latest_demand", "unreceived_pipeline_0", "unreceived_pipeline_1",
         "unreceived_pipeline_2", "unreceived_pipeline_3")


class _SupplyNetworkView:
    """The slice of `SupplyNetwork` that code outside the env touches (supplynetwork.py:37,114-171)."""

    def __init__(self, env):
        self._env = env
        self.internal_nodes = list(env.spec.node_names)
        self.agent_managed_facilities = [env.spec.node_names[k] for k in env.spec.agents]
        self.order_sequence = list(env.spec.order_sequence) + [env.spec.external_name]

    def get_state(self, node_name: str) -> dict:
        """The 7 numbers of node `node_name` as the reference's dict (observation of the current state,
        global view)."""
        env = self._env
        k = env.spec.node_names.index(node_name)
        full = env._full_observation()
        v = full[7 * k:7 * k + 7]
        cast = (lambda x: int(x)) if env.sim.dtype == "int32" else float
        return {key: cast(x) for key, x in zip(_KEYS, v)}


class BeerGameEnv:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, spec: ScenarioSpec, device="cuda", dtype="int32", action_space=None, seed=None,
                 return_dict: bool = False):
        self.spec = spec
        self.return_dict = return_dict        # envs.py:59,77-78: {agent: state dict} instead of the flat array
        self.sim = BatchedSupplyNetwork(spec, 1, device, dtype)
        self.max_episode_steps = spec.max_episode_steps
        self.global_observable = spec.global_observable
        self.action_space = action_space if action_space is not None else \
            spaces.Box(spec.action_low, spec.action_high, shape=(spec.n_agents,))
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(spec.obs_dim,))
        self.period, self.terminal = 0, False
        self._r64 = torch.zeros((1,), dtype=torch.float64, device=self.sim.device)
        self._global_sim_spec = None
        self.sn = _SupplyNetworkView(self)
        self.seed(seed)

    # the reference draws from numpy's GLOBAL generator (seed() there only seeds an unused one, Q9);
    # here an explicit seed gives a private RandomState, no seed uses the global one like the reference
    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed) if seed is not None else np.random
        return [seed]

    def reset(self):
        init, demand = _dem.seeded_episode_inputs_from_streams(self.spec, [self._rs])
        self.sim.load_episode(init, demand)
        obs = self.sim.reset()
        self.period, self.terminal = 0, False
        return self._format(obs.cpu().numpy()[0])

    def step(self, quantity):
        if self.terminal:
            raise ValueError("Cannot take action when the state is terminal.")
        q = np.asarray(quantity)
        if q.dtype == object or not (np.issubdtype(q.dtype, np.number) or np.issubdtype(q.dtype, np.bool_)):
            raise TypeError(f"order_quantity must be an instance of float, int or DemandGenerator, got {type(quantity)}")
        q = q.reshape(1, -1)
        if self.spec.rule == "a" and (q < 0).any():       # network_components.py:368-374 (the kernel clamps)
            warnings.warn(f"order quantity is {q.min()} but it should be non-negative. Quantity will be truncated to 0",
                          category=RuntimeWarning)
        if q.shape[1] != self.spec.n_agents:
            raise ValueError(f"expected {self.spec.n_agents} order quantities, got {q.shape[1]}")
        if self.spec.demand_pattern == "list" and self.period + 1 >= len(self.spec.demand_list):
            # the period advance reads demand[period + 1] (supplynetwork.py:266-268): past the end of a list pattern the
            # reference dies with numpy's IndexError (make_beer_game's 'deterministic_random' does so at its last step)
            size = len(self.spec.demand_list)
            raise IndexError(f"index {size} is out of bounds for axis 0 with size {size}")
        obs, _, done = self.sim.step(torch.as_tensor(q), reward64_out=self._r64)
        self.period, self.terminal = self.sim.period, done
        return self._format(obs.cpu().numpy()[0]), float(self._r64.item()), bool(done), {}

    def _format(self, flat: np.ndarray):
        if not self.return_dict:
            return flat
        cast = (lambda x: int(x)) if self.sim.dtype == "int32" else float
        names = [self.spec.node_names[k] for k in self.spec.obs_nodes]
        return {name: {key: cast(x) for key, x in zip(_KEYS, flat[7 * i:7 * i + 7])} for i, name in enumerate(names)}

    def _full_observation(self) -> np.ndarray:
        """Global observation (7 numbers per facility) of the current state, whatever this env's own observability."""
        if self.spec.global_observable:
            return self.sim.observe().cpu().numpy()[0]
        from dataclasses import replace
        import ctypes as C
        from . import _lib
        g = replace(self.spec, global_observable=True).to_c()
        out = torch.zeros((1, 7 * self.spec.n_facilities), dtype=torch.float32, device=self.sim.device)
        s = self.sim
        _lib.check(s.lib.sn_observe(C.byref(g), s.dtype_code, 1, s.ld, C.c_void_p(s.state.data_ptr()),
                                    C.c_void_p(out.data_ptr()), s._stream()))
        return out.cpu().numpy()[0]

    def close(self):
        pass

    def render(self, mode="human"):
        return None
