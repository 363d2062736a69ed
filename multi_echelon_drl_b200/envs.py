"""Env factories with the reference's names and arguments (envs.py:163-286, 289-414, 501-596),
returning a batched `SupplyNetVecEnv` instead of one gym env.

    ref:   env  = make_beer_game_uniform_multi_facility(agent_managed_facilities=[...], box_action_space=True, ...)
           venv = make_vec_env(lambda: env, n_env=2)
    here:  venv = make_beer_game_uniform_multi_facility(agent_managed_facilities=[...], box_action_space=True, ...,
                                                       n_envs=65536, device="cuda")

Extra keyword arguments (n_envs, device, dtype, seed, episode_source, output, use_graph) select the batch; every
reference argument keeps its meaning.  Without `n_envs` the factories return a single gym-style env
(`BeerGameEnv`, the duck type of the reference's InventoryManagementEnvMultiPlayer), exactly like the
reference; with `n_envs=N` they return the batched `SupplyNetVecEnv`.  `env_mode="test"` needs the demand sample files the reference
loads from ./data/deepbeerinventory/*.npy, which are not part of the reference repo either.
"""
from __future__ import annotations

from . import spaces
from . import spec as _spec
from .gym_env import BeerGameEnv
from .vec_env import SupplyNetVecEnv


def _apply_mode(sp, env_mode, data_path):
    """env_mode='test' replays demand rows from a sample file (envs.py:183-185, 309-311); the reference
    repository does not ship those .npy files, so the path must exist here too."""
    import os
    from dataclasses import replace
    if env_mode == "train":
        return sp
    if env_mode != "test":
        raise ValueError(env_mode)
    if not os.path.isfile(data_path):
        raise FileNotFoundError(f"env_mode='test' needs the demand sample file {data_path} (not part of the reference repo)")
    return replace(sp, demand_pattern="samples", demand_path=data_path)


def _build(sp, n_envs, device, dtype, seed, episode_source, output, box_action_space, return_dict=False,
           use_graph=False):
    if len(sp.agents) > 1 and not box_action_space:
        raise ValueError("length of agent_managed_facilities >= 1, only box_action_space is allowed. Please specify"
                         "box_action_space=True")                                            # envs.py:175-179
    if n_envs is None:             # the reference's return type: one gym-style env
        space = spaces.Box(sp.action_low, sp.action_high, shape=(sp.n_agents,)) if box_action_space else \
            spaces.Discrete(int(sp.action_high) + 1)
        return BeerGameEnv(sp, device, dtype or "int32", space, seed, return_dict)
    if return_dict:
        raise NotImplementedError("return_dict=True (per-agent dict observations) is offered by the single env only")
    return SupplyNetVecEnv(sp, n_envs, device, dtype or "float32", 0 if seed is None else seed, episode_source, output,
                           bool(box_action_space), use_graph=use_graph)


def make_beer_game_uniform_multi_facility(global_observable=False, agent_managed_facilities=None,
                                          max_episode_steps=100, return_dict=False, random_init=True,
                                          env_mode="train", box_action_space=False, *, n_envs=None, device="cuda",
                                          dtype=None, seed=None, episode_source="device", output="torch", use_graph=False):
    """The complex scenario (envs.py:289-414): demand U{0..8}, h=0.5, b=1, co-player targets 19/20/20/14."""
    sp = _spec.uniform_spec(agent_managed_facilities, global_observable, max_episode_steps, random_init)
    sp = _apply_mode(sp, env_mode, "./data/deepbeerinventory/demandTs0-9.npy")
    if sp.demand_pattern == "samples" and episode_source == "device":
        episode_source = "numpy_bulk"
    return _build(sp, n_envs, device, dtype, seed, episode_source, output, box_action_space, return_dict, use_graph)


def make_beer_game_normal_multi_facility(global_observable=False, agent_managed_facilities=None,
                                         max_episode_steps=100, return_dict=False, random_init=False,
                                         env_mode="train", box_action_space=False, *, n_envs=None, device="cuda",
                                         dtype=None, seed=None, episode_source="device", output="torch", use_graph=False):
    """The basic scenario (envs.py:163-286): demand round(max(N(10,2),0)), h=(1,.75,.5,.25), b=(10,0,0,0)."""
    sp = _spec.normal_spec(agent_managed_facilities, global_observable, max_episode_steps, random_init)
    sp = _apply_mode(sp, env_mode, "./data/deepbeerinventory/demandTs1-10-2.npy")
    if sp.demand_pattern == "samples" and episode_source == "device":
        episode_source = "numpy_bulk"
    return _build(sp, n_envs, device, dtype, seed, episode_source, output, box_action_space, return_dict, use_graph)


def make_beer_game(agent_managed_facilities=None, demand_type="classic_beer_game", max_period=35, return_dict=False,
                   random_init=False, seed=None, *, n_envs=None, device="cuda", dtype="int32",
                   episode_source="numpy_bulk", output="torch", use_graph=False):
    """The classic beer game (envs.py:501-596): demand 4,4,4,4,8,8,...; MultiDiscrete([17]*F) actions.
    `random_init` is accepted and ignored exactly as in the reference (never passed to Arc, Q16)."""
    if demand_type not in ("classic_beer_game", "normal", "uniform", "deterministic_random"):
        raise ValueError()
    import numpy as np
    if n_envs is None and seed:
        np.random.seed(seed)                         # envs.py:511-512
    if demand_type == "deterministic_random":
        # envs.py:519-520: ONE sequence of max_period values drawn from the global generator when the env is built and
        # replayed by every episode.  The period advance reads one value more than the list holds, so the reference's
        # episode ends in an IndexError at its last step; the single env reproduces that, the batched env is not offered.
        if n_envs is not None:
            raise NotImplementedError("demand_type='deterministic_random' cannot finish an episode in the reference "
                                      "(IndexError at the last step, envs.py:519-520): single env only")
        sp = _spec.classic_spec(agent_managed_facilities, max_period, "list",
                                (8 + 2 * np.random.randn(max_period)).astype(int))
    else:
        sp = _spec.classic_spec(agent_managed_facilities, max_period, demand_type)
    if n_envs is None:
        return BeerGameEnv(sp, device, dtype, spaces.MultiDiscrete([17] * sp.n_agents), None, return_dict)
    if return_dict:
        raise NotImplementedError("return_dict=True (per-agent dict observations) is offered by the single env only")
    return SupplyNetVecEnv(sp, n_envs, device, dtype, 0 if seed is None else seed, episode_source, output,
                           box_action_space=False, multi_discrete=True, use_graph=use_graph)


def make_env_from_config(network_config, agent_managed_facilities=None, global_observable=False, max_episode_steps=100,
                         return_dict=False, random_init=True, box_action_space=True, *, n_envs=None, device="cuda",
                         dtype=None, seed=None, episode_source="device", output="torch", use_graph=False):
    """What `SupplyNetwork.from_yaml` / `from_dict` + InventoryManagementEnvMultiPlayer were meant to give
    (supplynetwork.py:345-397; broken in the reference, Q1): an env for a network description - a path to a
    network_configs/*.yaml file or the dict read from one.  Serial chains retailer..manufacturer and distribution
    trees (one supplier per node) of up to four internal nodes; single-item (multi-item chains: multi_item.py)."""
    from .utils.graph import read_yaml
    cfg = read_yaml(network_config) if isinstance(network_config, str) else network_config
    specs = _spec.specs_from_config(cfg, agent_managed_facilities, global_observable, max_episode_steps,
                                    random_init=random_init)
    if len(specs) != 1:
        raise NotImplementedError("multi-item descriptions go through multi_item.MultiItemSupplyNetwork")
    return _build(specs[0], n_envs, device, dtype, seed, episode_source, output, box_action_space, return_dict, use_graph)
