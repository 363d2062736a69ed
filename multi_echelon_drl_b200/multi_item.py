"""Multi-item networks (BASELINE config 4, network_configs/multi_item.yaml).

The reference has no working multi-item behaviour: its multi_item.yaml is a TODO stub, `from_dict`
ignores `items` and crashes (supplynetwork.py:359-365), and list-valued costs would break `get_cost`
(supplynetwork.py:299).  Semantics defined here (parity unpinned): I items are I
independent copies of the chain with per-item cost vectors; the observation is the concatenation
of the items' observations ([N, I*obs_dim]), the action the concatenation of the items' order
columns ([N, I*n_agents]) and the reward the sum over items.  Each item is one
`BatchedSupplyNetwork` (own spec, own state rows) driven by the same kernels.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .simulator import BatchedSupplyNetwork
from .spec import ScenarioSpec


class MultiItemSupplyNetwork:
    def __init__(self, specs: Sequence[ScenarioSpec], n_envs: int, device="cuda", dtype="int32"):
        if len(specs) < 1:
            raise ValueError("need at least one item")
        if len({(s.agents, s.global_observable, s.max_episode_steps) for s in specs}) != 1:
            raise ValueError("all items must share agents, observability and episode length")
        self.items: List[BatchedSupplyNetwork] = [BatchedSupplyNetwork(s, n_envs, device, dtype) for s in specs]
        it = self.items[0]
        self.n_items, self.n_envs, self.device = len(specs), n_envs, it.device
        self.item_obs_dim, self.item_agents = it.obs_dim, it.n_agents
        self.obs_dim, self.n_agents = self.n_items * it.obs_dim, self.n_items * it.n_agents
        self.T = it.T
        self.obs = torch.zeros((n_envs, self.obs_dim), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((n_envs,), dtype=torch.float32, device=self.device)
        self._r64 = [torch.zeros((n_envs,), dtype=torch.float64, device=self.device) for _ in specs]

    @property
    def period(self):
        return self.items[0].period

    @property
    def terminal(self):
        return self.items[0].terminal

    def load_episode(self, init, demand):
        """init [n, I, 19], demand [n, I, >=T+1]: per-item streams."""
        init, demand = torch.as_tensor(init), torch.as_tensor(demand)
        for i, it in enumerate(self.items):
            it.load_episode(init[:, i], demand[:, i])

    def _gather_obs(self):
        d = self.item_obs_dim
        for i, it in enumerate(self.items):
            self.obs[:, i * d:(i + 1) * d] = it.obs
        return self.obs

    def reset(self):
        for it in self.items:
            it.reset()
        return self._gather_obs()

    def step(self, actions):
        actions = torch.as_tensor(actions)
        if tuple(actions.shape) != (self.n_envs, self.n_agents):
            raise ValueError(f"actions must have shape {(self.n_envs, self.n_agents)}, got {tuple(actions.shape)}")
        a = self.item_agents
        done = False
        for i, it in enumerate(self.items):
            _, _, done = it.step(actions[:, i * a:(i + 1) * a], reward64_out=self._r64[i])
        total = self._r64[0].clone()
        for r in self._r64[1:]:
            total += r
        self.reward.copy_(total)                 # summed in double, rounded once like the single-item reward
        return self._gather_obs(), self.reward, done

    @property
    def ep_return(self):
        tot = self.items[0].ep_return.clone()
        for it in self.items[1:]:
            tot += it.ep_return
        return tot
