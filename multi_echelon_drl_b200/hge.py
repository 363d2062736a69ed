"""Heuristic-guided exploration: the action-selection branch of the reference's `HgeTD3.predict`
(hge.py:144-172) for device batches.  The RL algorithm itself (SB3 TD3) is out of scope; this is the
contract between it, the heuristic and the VecEnv:

    with probability hge_rate**2 per *call* (one coin for the whole batch, hge.py:157, Q8) the action is
    heuristic.get_order_quantity(original_obs)   (un-normalised observations, hge.py:125,164)
    otherwise it is policy(observation).

`per_env=True` flips one coin per env instead (not what the reference does; offered for large batches).
`HgeRateSchedule` is the linear decay of utils/callbacks.py:5-23.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


def hge_predict(policy: Callable[[torch.Tensor], torch.Tensor], heuristic, observation: torch.Tensor,
                original_obs: torch.Tensor, hge_rate: float, deterministic: bool = False, per_env: bool = False,
                rng: Optional[np.random.RandomState] = None, generator: Optional[torch.Generator] = None):
    """Returns (actions [n, n_agents] on the device, used_heuristic: bool or bool tensor [n])."""
    p = float(hge_rate) ** 2
    if deterministic or p <= 0.0:
        return policy(observation), False
    if not per_env:
        coin = (rng.rand() if rng is not None else np.random.rand()) < p
        if coin:
            return heuristic.get_order_quantity(original_obs), True
        return policy(observation), False
    mask = torch.rand(observation.shape[0], device=observation.device, generator=generator) < p
    a_pol = policy(observation)
    a_heu = heuristic.get_order_quantity(original_obs).to(a_pol.dtype)
    return torch.where(mask[:, None], a_heu, a_pol), mask


class HgeRateSchedule:
    """HgeRateCallback (utils/callbacks.py:5-23): rate = max(0, start * (1 - timesteps/100/decay_periods))."""

    def __init__(self, hge_rate_at_start: float, decay_periods: int = 1000):
        self.start, self.decay_periods = float(hge_rate_at_start), int(decay_periods)

    def __call__(self, num_timesteps: int) -> float:
        return max(0.0, self.start * (1.0 - num_timesteps / 100 / self.decay_periods))
