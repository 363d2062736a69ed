"""ctypes binding of the C ABI in include/supplynet.h (libsupplynet.so, built in-tree).

There is deliberately no fallback: if the shared library is missing or a CUDA call fails the
caller gets an exception, never a CPU result.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SUPPLYNET_LIB", os.path.join(_HERE, "csrc", "libsupplynet.so"))   # override: A/B builds

SN_I32, SN_F32, SN_F64 = 0, 1, 2
SN_RULE_A, SN_RULE_D_PLUS_A = 0, 1
SN_TOPO_CHAIN, SN_TOPO_TREE = 0, 1
SN_POLICY_ACTIONS, SN_POLICY_BASE_STOCK = 0, 1
SN_STATE_WORDS, SN_INIT_WORDS, SN_STATS_WORDS = 40, 19, 16
SN_OK, SN_ERR_INVALID, SN_ERR_UNSUPPORTED, SN_ERR_TERMINAL, SN_ERR_CUDA, SN_ERR_TYPE = 0, -1, -2, -3, -4, -5
ABI_VERSION = 11
SN_CTL_WORDS, SN_CTL_PERIOD, SN_CTL_TICKET, SN_CTL_FLAGS, SN_CTL_SEQ = 8, 0, 1, 2, 4
SN_VN_ONEPASS_SCRATCH_DOUBLES = 256 * 128
SN_FLAG_STEP_AFTER_TERMINAL, SN_FLAG_QTY_RANGE = 1, 2
SN_QTY_LIMIT_LOG2 = 26
SN_SRC_I32, SN_SRC_I64, SN_SRC_F32, SN_SRC_F64 = 0, 1, 2, 3


MAXF = 8            # SN_MAX_FACILITIES: size of the per-facility arrays (chains use the first four)


class SnSpec(C.Structure):
    _fields_ = [
        ("n_facilities", C.c_int32),
        ("info_leadtime", C.c_int32 * MAXF),
        ("ship_leadtime", C.c_int32 * MAXF),
        ("n_agents", C.c_int32),
        ("agents", C.c_int32 * MAXF),
        ("global_observable", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("rule", C.c_int32),
        ("holding_cost", C.c_double * MAXF),
        ("backorder_cost", C.c_double * MAXF),
        ("coplayer_target", C.c_double * MAXF),
        ("coplayer_lb", C.c_double),
        ("coplayer_ub", C.c_double),
        ("dpa_offset", C.c_double),
        ("dpa_lb", C.c_double),
        ("dpa_ub", C.c_double),
        ("n_items", C.c_int32),
        ("topology", C.c_int32),
        ("item_holding_cost", (C.c_double * MAXF) * 3),
        ("item_backorder_cost", (C.c_double * MAXF) * 3),
        ("supplier", C.c_int32 * MAXF),
        ("demand_source", C.c_int32 * MAXF),
        ("order_pos", C.c_int32 * MAXF),
        ("customer_rank", C.c_int32 * MAXF),
    ]


class SnPolicy(C.Structure):
    _fields_ = [
        ("target", C.c_double * MAXF),
        ("lb", C.c_double),
        ("ub", C.c_double),
        ("rule", C.c_int32),
        ("row0_broadcast", C.c_int32),
    ]


class SnStepIo(C.Structure):
    """sn_step_io of include/supplynet.h: buffers of one period step (device pointers as integers / None)."""
    _fields_ = [
        ("state", C.c_void_p), ("actions", C.c_void_p), ("demand", C.c_void_p),
        ("period", C.c_int32), ("reserved", C.c_int32),
        ("ctl", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p), ("env_reward", C.c_void_p),
        ("reward64", C.c_void_p), ("cost", C.c_void_p), ("ep_return", C.c_void_p), ("stats", C.c_void_p),
        ("vn_scratch", C.c_void_p), ("vn_obs_rms", C.c_void_p), ("vn_ret_rms", C.c_void_p),
        ("vn_returns", C.c_void_p), ("vn_gamma", C.c_double),
        ("vn_nobs", C.c_void_p), ("vn_nrew", C.c_void_p), ("vn_clip_obs", C.c_float), ("vn_clip_reward", C.c_float),
        ("vn_epsilon", C.c_double),
    ]


class SnReplay(C.Structure):
    """sn_replay of include/supplynet.h: the device replay buffer's arrays and its counters in device memory."""
    _fields_ = [
        ("obs", C.c_void_p), ("next_obs", C.c_void_p), ("act", C.c_void_p), ("rew", C.c_void_p), ("done", C.c_void_p),
        ("capacity", C.c_int64), ("obs_dim", C.c_int32), ("act_dim", C.c_int32),
        ("pos", C.c_void_p), ("size", C.c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/supplynet.h one to one
_P, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
SIGNATURES = {
    "sn_abi_version": (_I32, []),
    "sn_last_error": (C.c_char_p, []),
    "sn_obs_dim": (_I32, [C.POINTER(SnSpec)]),
    "sn_spec_validate": (_I32, [C.POINTER(SnSpec), _I32]),
    "sn_state_words": (_I32, [C.POINTER(SnSpec)]),
    "sn_init_words": (_I32, [C.POINTER(SnSpec)]),
    "sn_reset": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, _P, _P, _P, _P, _P]),
    "sn_step": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sn_step_ex": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, C.POINTER(SnStepIo), _P]),
    "sn_observe": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, _P, _P, _P]),
    "sn_heuristic": (_I32, [C.POINTER(SnPolicy), _I32, _I64, _P, _P, _P]),
    "sn_rollout": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, _I32, _I32, _P, _P, _I32, _P, C.POINTER(SnPolicy),
                          _P, _P, _P, _P, _P, _P, _P]),
    "sn_vecnorm_update": (_I32, [_I64, _I32, _P, _P, C.c_double, _P, _P, _P, _P, _P, _P, _P]),
    "sn_vecnorm_apply": (_I32, [_I64, _I32, _P, _P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_double, _I32, _P, _P]),
    "sn_fill_uniform": (_I32, [_I32, _P, _I64, _I32, _I32, C.c_uint64, C.c_uint64, _P]),
    "sn_fill_normal_demand": (_I32, [_I32, _P, _I64, C.c_double, C.c_double, C.c_uint64, C.c_uint64, _P]),
    "sn_fill_uniform_at": (_I32, [_I32, _P, _I64, _I32, _I32, C.c_uint64, _P, C.c_uint64, _P]),
    "sn_fill_normal_demand_at": (_I32, [_I32, _P, _I64, C.c_double, C.c_double, C.c_uint64, _P, C.c_uint64, _P]),
    "sn_counter_add": (_I32, [_P, C.c_uint64, _P]),
    "sn_pack_rows": (_I32, [_I32, _P, _I64, _I32, _I64, _I32, _P, _I32, _I64, _I64, _P]),
    "sn_replay_add": (_I32, [C.POINTER(SnReplay), _I64, _P, _P, _P, _P, _P, _P, _P]),
    "sn_replay_sample": (_I32, [C.POINTER(SnReplay), _I64, C.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    "sn_mail_bytes": (_I64, [_I32]),
    "sn_mail_alloc": (_I32, [_I32, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "sn_mail_open": (_I32, [C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "sn_mail_close": (_I32, [_P]),
    "sn_mail_free": (_I32, [_P]),
    "sn_enable_peer_access": (_I32, [_I32, _I32]),
    "sn_stats_allreduce": (_I32, [_P, _P, _P, _I32, _I32, C.c_uint32, _P]),
    "sn_episode": (_I32, [C.POINTER(SnSpec), _I32, _I64, _I64, _I32, _P, _P, _I32, _P, C.POINTER(SnPolicy),
                          _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


class SupplyNetError(RuntimeError):
    pass


def load():
    """Load libsupplynet.so (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C multi_echelon_drl_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    if lib.sn_abi_version() != ABI_VERSION:
        raise ImportError(f"libsupplynet ABI {lib.sn_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, lib=None):
    """Map status codes to the exceptions the reference raises at the same places.  `lib`: the library that
    returned `rc` when it is not the main one (a specialised build keeps its own error text)."""
    if rc >= 0:
        return rc
    msg = (lib if lib is not None else load()).sn_last_error().decode()
    if rc == SN_ERR_TERMINAL or rc == SN_ERR_INVALID:
        raise ValueError(msg)                      # envs.py:88-89
    if rc == SN_ERR_TYPE:
        raise TypeError(msg)                       # network_components.py:26-29
    if rc == SN_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SupplyNetError(msg)
