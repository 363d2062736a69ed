"""Env ids of the reference registry (register_envs.py:72-165), resolvable without gym.

    ref:   register_envs(); env = make_vec_env(lambda: gym.make(env_id), n_env)
    here:  venv = make(env_id, n_envs=..., device="cuda")            # or make_vec_env(env_id, n_envs)

38 ids: BeerGame{Uniform,Normal}{Retailer,Wholesaler,Distributor,Manufacturer}[Discrete][FullInfo]-v0,
BeerGame{Uniform,Normal}MultiFacilityFullInfo-v0, BeerGame{Uniform,Normal}MultiFacilityDPlusA-v0,
BeerGameUniformRetailer[Discrete]DPlusA-v0; all with max_episode_steps=100 and random_init=True.

Quirk kept for literal parity: the reference passes `box_action_space=~discrete`, and `~True == -2` is
truthy, so every "Discrete" id of the role loop still gets a Box action space (Q3); only
BeerGameUniformRetailerDiscreteDPlusA-v0 (built with box_action_space=False) is really Discrete(17).
"""
from __future__ import annotations

from typing import Callable, Dict

from .envs import make_beer_game_normal_multi_facility, make_beer_game_uniform_multi_facility
from .utils.wrappers import wrap_action_d_plus_a

ALL = ["retailer", "wholesaler", "distributor", "manufacturer"]
_REGISTRY: Dict[str, Callable] = {}


def _factory(kind, role, discrete, global_observable, dplusa=None, box=None):
    maker = make_beer_game_uniform_multi_facility if kind == "uniform" else make_beer_game_normal_multi_facility

    def func(**batch_kwargs):
        env = maker(agent_managed_facilities=role, box_action_space=(~int(discrete) if box is None else box),
                    random_init=True, global_observable=global_observable, max_episode_steps=100, **batch_kwargs)
        if dplusa is not None:
            env = wrap_action_d_plus_a(env, offset=dplusa[0], lb=dplusa[1], ub=dplusa[2])
        return env

    return func


def register_envs():
    r = _REGISTRY
    r["BeerGameUniformMultiFacilityDPlusA-v0"] = _factory("uniform", ALL, False, False, (-8, 0, 16), box=True)
    r["BeerGameNormalMultiFacilityDPlusA-v0"] = _factory("normal", ALL, False, False, (-10, 0, 20), box=True)
    r["BeerGameUniformRetailerDPlusA-v0"] = _factory("uniform", ["retailer"], False, False, (-8, 0, 16), box=True)
    r["BeerGameUniformRetailerDiscreteDPlusA-v0"] = _factory("uniform", ["retailer"], True, False, (-8, 0, 16), box=False)
    r["BeerGameUniformMultiFacilityFullInfo-v0"] = _factory("uniform", ALL, False, True)
    r["BeerGameNormalMultiFacilityFullInfo-v0"] = _factory("normal", ALL, False, True)
    for key in ("Retailer", "Wholesaler", "Distributor", "Manufacturer"):
        role = [key.lower()]
        for kind, tag in (("uniform", "Uniform"), ("normal", "Normal")):
            r[f"BeerGame{tag}{key}-v0"] = _factory(kind, role, False, False)
            r[f"BeerGame{tag}{key}Discrete-v0"] = _factory(kind, role, True, False)
            r[f"BeerGame{tag}{key}FullInfo-v0"] = _factory(kind, role, False, True)
            r[f"BeerGame{tag}{key}DiscreteFullInfo-v0"] = _factory(kind, role, True, True)
    return sorted(r)


def registered_ids():
    if not _REGISTRY:
        register_envs()
    return sorted(_REGISTRY)


def make(env_id: str, n_envs=None, **batch_kwargs):
    """`make(id)` -> one gym-style env like `gym.make(id)` in the reference (episode length 100 =
    the registry's max_episode_steps, so no TimeLimit wrapper is needed); `make(id, n_envs=N)` -> the
    batched VecEnv."""
    if not _REGISTRY:
        register_envs()
    if env_id not in _REGISTRY:
        raise KeyError(f"No registered env with id: {env_id}")
    return _REGISTRY[env_id](n_envs=n_envs, **batch_kwargs)


def make_vec_env(env_id: str, n_envs: int = 1, **batch_kwargs):
    """Counterpart of stable_baselines3.common.env_util.make_vec_env for registry ids."""
    return make(env_id, n_envs, **batch_kwargs)
