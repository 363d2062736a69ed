"""Base-stock heuristic with the reference's interface (utils/heuristics.py:5-75), evaluated by the
`sn_heuristic` CUDA kernel for device batches.

    bsp = BaseStockPolicy(target_levels=[19, 20, 20, 14], array_index=..., state_dim_per_facility=7,
                          lb=0, ub=16, rule='a')
    actions = bsp.get_order_quantity(original_obs)      # torch CUDA tensor [n, 7F] -> [n, F]

Differences from the reference, on purpose:
  * a batch is evaluated per env.  The reference flattens the batch and reads only row 0
    (utils/heuristics.py:44-52); `row0_broadcast=True` reproduces that literally.
  * the dict path (per-agent state dicts, used by the reference's co-players inside the env) goes through
    the same kernel (there is no CPU evaluation anywhere); inside the batched env the co-players' policy is
    part of the step kernel.
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
            # reference dict path (utils/heuristics.py:60-75): one entry per agent -> shape (n_agents, n_targets),
            # every target applied to that agent's own state.  Evaluated by the same CUDA kernel: the agent's 7
            # numbers are replicated once per target column.
            keys = ("on_hand", "unfilled_demand", "latest_demand", "unreceived_pipeline_0", "unreceived_pipeline_1",
                    "unreceived_pipeline_2", "unreceived_pipeline_3")
            m = self.target_levels.shape[0]
            if m > 8:
                raise NotImplementedError("the kernel evaluates at most 8 targets per state")
            rows = np.array([[st[k] for k in keys] * m for st in observation.values()], dtype=np.float32)
            out = heuristic_actions(self.params, torch.as_tensor(rows).cuda())
            return out.cpu().numpy().astype(np.float64)
        if isinstance(observation, np.ndarray):
            observation = torch.as_tensor(observation, dtype=torch.float32).cuda()
            return heuristic_actions(self.params, observation).cpu().numpy().astype(np.float64)
        if isinstance(observation, torch.Tensor):
            return heuristic_actions(self.params, observation)
        raise TypeError
