"""ctypes binding of oracle/beergame_c.c (TEST INFRASTRUCTURE / CPU baseline, not product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbeergame_oracle.so")


class BgSpec(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("agents", C.c_int32 * 4), ("global_observable", C.c_int32),
                ("max_episode_steps", C.c_int32), ("rule", C.c_int32), ("holding_cost", C.c_double * 4),
                ("backorder_cost", C.c_double * 4), ("coplayer_target", C.c_int64 * 4), ("coplayer_lb", C.c_int64),
                ("coplayer_ub", C.c_int64), ("dpa_offset", C.c_int64), ("dpa_lb", C.c_int64), ("dpa_ub", C.c_int64)]


class BgPolicy(C.Structure):
    _fields_ = [("target", C.c_int64 * 4), ("lb", C.c_int64), ("ub", C.c_int64), ("rule", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "beergame_c.c")):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = C.CDLL(_SO)
        lib.bg_rollout.restype = C.c_int64
        lib.bg_rollout.argtypes = [C.POINTER(BgSpec), C.c_int64, C.c_int32] + [C.c_void_p] * 3 + \
                                  [C.POINTER(BgPolicy), C.c_int32] + [C.c_void_p] * 4
        lib.bg_num_threads.restype = C.c_int32
        _lib = lib
    return _lib


def make_spec(ospec) -> BgSpec:
    """From an oracle.beergame_np.OracleSpec."""
    s = BgSpec()
    s.n_agents = len(ospec.agents)
    for i, a in enumerate(ospec.agents):
        s.agents[i] = a
    s.global_observable = int(ospec.global_observable)
    s.max_episode_steps = ospec.max_episode_steps
    s.rule = 1 if ospec.rule == "d+a" else 0
    for k in range(4):
        s.holding_cost[k] = ospec.holding_cost[k]
        s.backorder_cost[k] = ospec.backorder_cost[k]
        s.coplayer_target[k] = int(ospec.coplayer_target[k])
    s.coplayer_lb, s.coplayer_ub = int(ospec.coplayer_lb), int(ospec.coplayer_ub)
    s.dpa_offset, s.dpa_lb, s.dpa_ub = int(ospec.dpa_offset), int(ospec.dpa_lb), int(ospec.dpa_ub)
    return s


def rollout(ospec, init, demand, actions=None, policy=None, n_steps=None, n_threads=1, want_obs=True,
            want_reward=True):
    """init [n,19], demand [n,T+3], actions [steps,n,na] ints or policy=(targets, lb, ub, rule)."""
    lib = load()
    s = make_spec(ospec)
    init = np.ascontiguousarray(init, dtype=np.int32)
    demand = np.ascontiguousarray(demand, dtype=np.int32)
    n = init.shape[0]
    assert demand.shape == (n, ospec.max_episode_steps + 3), demand.shape
    na = len(ospec.agents)
    od = 7 * (4 if ospec.global_observable else na)
    pol = None
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        n_steps = actions.shape[0] if n_steps is None else n_steps
        assert actions.shape[1:] == (n, na)
    else:
        tg, lb, ub, rule = policy
        pol = BgPolicy()
        for i in range(na):
            pol.target[i] = int(tg[i])
        pol.lb, pol.ub, pol.rule = int(lb), int(ub), 1 if rule == "d+a" else 0
        n_steps = ospec.max_episode_steps if n_steps is None else n_steps
    obs0 = np.zeros((n, od), np.float32)
    obs = np.zeros((n_steps, n, od), np.float32) if want_obs else None
    rew = np.zeros((n_steps, n), np.float64) if want_reward else None
    ret = np.zeros((n,), np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    done = lib.bg_rollout(C.byref(s), n, n_steps, p(init), p(demand), p(actions), C.byref(pol) if pol else None,
                          n_threads, p(obs0), p(obs), p(rew), p(ret))
    assert done == n * n_steps
    return dict(obs0=obs0, obs=obs, rew=rew, ep_return=ret)
