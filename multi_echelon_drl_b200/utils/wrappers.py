"""The d+a ordering rule (reference utils/wrappers.py:6-31) for the batched env.

The reference monkey-patches env.step so that the action `a` is turned into the order
clip(latest_demand + a + offset, lb, ub).  Here the rule is a field of the scenario spec and is
applied inside the step / rollout kernels; `wrap_action_d_plus_a` keeps the reference's name and
arguments and returns the same env object with the rule switched on."""
from __future__ import annotations


def wrap_action_d_plus_a(env, offset=-8, lb: int = 0, ub: int = 16):
    from .. import _lib
    import ctypes as C
    base = env
    while hasattr(base, "venv"):           # reach the SupplyNetVecEnv through wrappers such as VecNormalize
        base = base.venv
    sim = base.sim
    spec = base.spec.with_rule("d+a", offset, lb, ub)
    c_spec = spec.to_c()
    _lib.check(sim.lib.sn_spec_validate(C.byref(c_spec), sim.dtype_code))
    base.spec = spec
    sim.spec, sim._c_spec = spec, c_spec
    if getattr(base, "_graph", None) is not None:      # a step graph captured before the switch has the old rule baked in
        base._graph = None
    return env
