"""B200-native batched supply-network (beer game) simulator.

Drop-in for the data-parallel hot path of qihuazhong/multi-echelon-drl: stepping many
independent multi-echelon inventory environments in lockstep and evaluating the base-stock /
HGE heuristic, behind the reference's env / VecEnv / heuristic interfaces.
"""
from .spec import ScenarioSpec, uniform_spec, normal_spec, classic_spec, NODE_NAMES  # noqa: F401

__all__ = ["ScenarioSpec", "uniform_spec", "normal_spec", "classic_spec", "NODE_NAMES"]
