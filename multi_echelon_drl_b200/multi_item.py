"""Multi-item networks (BASELINE config 4, network_configs/multi_item.yaml).

The reference has no working multi-item behaviour: its multi_item.yaml is a TODO stub, `from_dict`
ignores `items` and crashes (supplynetwork.py:359-365), and list-valued costs would break `get_cost`
(supplynetwork.py:299).  Semantics defined here (parity unpinned): I items are I
independent copies of the chain with per-item cost vectors; the observation of an env is the
concatenation of its items' observations ([N, I*obs_dim]), the action the concatenation of the items'
order columns ([N, I*n_agents]) and the reward the sum over items.

Native layout: ONE batch of N*I chains, chain c = env*I + item (item-minor).  The kernels pick the cost
vector by `c % I` (sn_spec.n_items / item_*_cost), so the row-major tensors they already read and
write ARE the concatenated ones: obs [N*I, obs_dim] viewed as [N, I*obs_dim], actions likewise.  No
gather/scatter pass exists; only the reward is reduced over the I adjacent chains of an env.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Sequence

import torch

from .simulator import BatchedSupplyNetwork
from .spec import ScenarioSpec


def combine_item_specs(specs: Sequence[ScenarioSpec]) -> ScenarioSpec:
    if not 1 <= len(specs) <= 4:
        raise ValueError("1..4 items are supported")
    base = specs[0]
    for s in specs[1:]:
        if replace(s, name=base.name, holding_cost=base.holding_cost, backorder_cost=base.backorder_cost) != base:
            raise ValueError("items may differ in holding/backorder cost only")
    return replace(base, extra_item_costs=tuple((tuple(s.holding_cost), tuple(s.backorder_cost)) for s in specs[1:]))


class MultiItemSupplyNetwork:
    def __init__(self, specs: Sequence[ScenarioSpec], n_envs: int, device="cuda", dtype="int32", use_graph: bool = False):
        self.spec = combine_item_specs(specs)
        self.n_items, self.n_envs = len(specs), int(n_envs)
        self.sim = BatchedSupplyNetwork(self.spec, self.n_envs * self.n_items, device, dtype)
        s = self.sim
        self.device, self.T = s.device, s.T
        self.item_obs_dim, self.item_agents = s.obs_dim, s.n_agents
        self.obs_dim, self.n_agents = self.n_items * s.obs_dim, self.n_items * s.n_agents
        self.obs = s.obs.view(self.n_envs, self.obs_dim)                       # same memory, concatenated view
        self.reward = torch.zeros((self.n_envs,), dtype=torch.float32, device=self.device)
        self._fused_reward = self.n_items in (2, 4)         # in-kernel reduction over the adjacent item chains
        self._r64 = None if self._fused_reward else torch.zeros((self.n_envs * self.n_items,), dtype=torch.float64,
                                                                device=self.device)
        self.use_graph = bool(use_graph) and self._fused_reward
        self._graph = None
        self._act = torch.zeros((self.n_envs * self.n_items, self.item_agents), dtype=s.tdtype, device=self.device)

    @property
    def period(self):
        return self.sim.period

    @property
    def terminal(self):
        return self.sim.terminal

    def load_episode(self, init, demand):
        """init [n, I, 19], demand [n, I, >=T+1]: per-item streams."""
        init, demand = torch.as_tensor(init), torch.as_tensor(demand)
        n, i = self.n_envs, self.n_items
        self.sim.load_episode(init.reshape(n * i, init.shape[-1]), demand.reshape(n * i, demand.shape[-1]))

    def reset(self):
        self.sim.reset()
        return self.obs

    def _capture(self):
        sim = self.sim
        sim._graph_period_valid = True
        sim.ctl[0:1].fill_(sim.period)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sim.launch_step_device_period(self._act, sim.obs, sim.reward, env_reward=self.reward)
        # capture does not execute: state and counter are untouched
        self._graph = g

    def step(self, actions):
        """actions [n, I*n_agents] (item-major columns) -> (obs [n, I*obs_dim], reward [n], done)."""
        actions = torch.as_tensor(actions)
        if tuple(actions.shape) != (self.n_envs, self.n_agents):
            raise ValueError(f"actions must have shape {(self.n_envs, self.n_agents)}, got {tuple(actions.shape)}")
        a = actions.reshape(self.n_envs * self.n_items, self.item_agents)     # a view when contiguous
        sim = self.sim
        if self.use_graph and not sim.terminal and sim.period + 1 < sim.T:
            if a.data_ptr() != self._act.data_ptr():
                self._act.copy_(sim._check_actions(a))
            if self._graph is None:
                self._capture()
            self._graph.replay()
            sim.advance_host_period()
            return self.obs, self.reward, sim.terminal
        if self._fused_reward:
            _, _, done = sim.step(a, env_reward_out=self.reward)
        else:
            _, _, done = sim.step(a, reward64_out=self._r64)
            # summed over the items of an env in double, rounded once like the single-item reward
            self.reward.copy_(self._r64.view(self.n_envs, self.n_items).sum(1))
        return self.obs, self.reward, done

    @property
    def action_buffer(self):
        """[n, I*n_agents] view of the static action buffer of the graph-replayed step: write actions here to
        skip the copy."""
        return self._act.view(self.n_envs, self.n_agents)

    @property
    def ep_return(self):
        return self.sim.ep_return.view(self.n_envs, self.n_items).sum(1)
