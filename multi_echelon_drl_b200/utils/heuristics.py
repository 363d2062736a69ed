"""Base-stock heuristic with the reference's interface (utils/heuristics.py:5-75), evaluated by the
`sn_heuristic` CUDA kernel for device batches.

    bsp = BaseStockPolicy(target_levels=[19, 20, 20, 14], array_index=..., state_dim_per_facility=7,
                          lb=0, ub=16, rule='a')
    actions = bsp.get_order_quantity(original_obs)      # torch CUDA tensor [n, 7F] -> [n, F]

Differences from the reference, on purpose:
  * a batch is evaluated per env.  The reference flattens the batch and reads only row 0
    (utils/heuristics.py:44-52); `row0_broadcast=True` reproduces that literally.
  * the dict path (per-agent state dicts, used by the reference's co-players inside the env) is kept
    as plain Python for API parity; inside the batched env the co-players' policy runs in the kernel.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

from ..simulator import BasePolicyParams, heuristic_actions

_DEFAULT_INDEX = {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1, "latest_demand": 2}


class InventoryPolicy:
    def get_order_quantity(self, observation):
        raise NotImplementedError


class BaseStockPolicy(InventoryPolicy):
    def __init__(self, target_levels: Union[int, float, List, np.ndarray], array_index: Optional[dict] = None,
                 state_dim_per_facility=7, lb=0, ub=16, rule: str = "a", row0_broadcast: bool = False):
        self.target_levels = np.asarray(target_levels).reshape(-1)
        self.array_index = array_index
        self.state_dim_per_facility = state_dim_per_facility
        self.lb, self.ub, self.rule = lb, ub, rule
        if array_index is not None and (state_dim_per_facility != 7 or
                                        {k: list(np.atleast_1d(v)) for k, v in array_index.items()} !=
                                        {k: list(np.atleast_1d(v)) for k, v in _DEFAULT_INDEX.items()}):
            raise NotImplementedError("the kernel implements the reference's observation layout "
                                      "(on_hand 0, unfilled 1, latest 2, pipeline 3..6; 7 per facility)")
        self.params = BasePolicyParams(self.target_levels, lb, ub, rule, row0_broadcast)

    def get_order_quantity(self, observation):
        if isinstance(observation, dict):
            # reference dict path (utils/heuristics.py:60-75): one entry per agent -> shape (n_agents, n_targets)
            out = []
            for _, st in observation.items():
                on_order = sum(st[f"unreceived_pipeline_{i}"] for i in range(4))
                q = self.target_levels - (st["on_hand"] + on_order - st["unfilled_demand"])
                if self.rule == "d+a":
                    q = q - st["latest_demand"]
                out.append(q)
            return np.clip(out, self.lb, self.ub)
        if isinstance(observation, np.ndarray):
            observation = torch.as_tensor(observation, dtype=torch.float32).cuda()
            return heuristic_actions(self.params, observation).cpu().numpy().astype(np.float64)
        if isinstance(observation, torch.Tensor):
            return heuristic_actions(self.params, observation)
        raise TypeError
