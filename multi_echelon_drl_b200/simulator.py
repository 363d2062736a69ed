"""BatchedSupplyNetwork: N independent serial beer games stepped in lockstep on one B200.

This is the device-side stand-in for N instances of the reference's `SupplyNetwork` wrapped in
`InventoryManagementEnvMultiPlayer` (supplynetwork.py:10-343, envs.py:29-117).  It owns the
structure-of-arrays state tensor and the per-episode demand table and drives the CUDA kernels
through the C ABI (include/supplynet.h).  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .spec import ScenarioSpec

_DTYPES = {
    "int32": (_lib.SN_I32, torch.int32), "i32": (_lib.SN_I32, torch.int32),
    "float32": (_lib.SN_F32, torch.float32), "f32": (_lib.SN_F32, torch.float32),
    "float64": (_lib.SN_F64, torch.float64), "f64": (_lib.SN_F64, torch.float64),
}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class BasePolicyParams:
    """Parameters of a base-stock policy (utils/heuristics.py:13-33) in C form."""

    def __init__(self, target_levels, lb=0, ub=16, rule="a", row0_broadcast=False):
        tl = np.asarray(target_levels, dtype=np.float64).reshape(-1)
        if not 1 <= tl.shape[0] <= _lib.MAXF:
            raise ValueError(f"target_levels must have 1..{_lib.MAXF} entries")
        if rule not in ("a", "d+a"):
            raise ValueError(f"rule must be 'a' or 'd+a', got {rule!r}")
        self.n_cols = tl.shape[0]
        self.c = _lib.SnPolicy()
        for i in range(_lib.MAXF):
            self.c.target[i] = float(tl[i]) if i < self.n_cols else 0.0
        self.c.lb, self.c.ub = float(lb), float(ub)
        self.c.rule = _lib.SN_RULE_D_PLUS_A if rule == "d+a" else _lib.SN_RULE_A
        self.c.row0_broadcast = int(bool(row0_broadcast))


class BatchedSupplyNetwork:
    def __init__(self, spec: ScenarioSpec, n_envs: int, device="cuda", dtype="int32"):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedSupplyNetwork needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self.spec = spec
        self.n_envs = int(n_envs)
        if self.n_envs < 1:
            raise ValueError("n_envs must be >= 1")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedSupplyNetwork needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype_code, self.tdtype = _DTYPES[dtype]
        self.dtype = dtype
        self.ld = (self.n_envs + 31) // 32 * 32            # rows 128-byte aligned
        self.T = spec.max_episode_steps
        self._c_spec = spec.to_c()
        _lib.check(self.lib.sn_spec_validate(C.byref(self._c_spec), self.dtype_code))
        self.obs_dim = spec.obs_dim
        self.n_agents = spec.n_agents
        dev, n, ld = self.device, self.n_envs, self.ld
        self.state_words = self.lib.sn_state_words(C.byref(self._c_spec))      # 40 (beer-game lead times) or 57
        self.init_words = self.lib.sn_init_words(C.byref(self._c_spec))        # 19 or 36
        self.state = torch.zeros((self.state_words, ld), dtype=self.tdtype, device=dev)
        self.init = torch.zeros((self.init_words, ld), dtype=self.tdtype, device=dev)
        # chains: demand[t] is the [ld] row of period t; trees: [n_sources][ld] rows per period (sources in node order)
        self.n_sources = spec.n_sources
        self.demand = torch.zeros((self.T + 3, self.n_sources, ld) if spec.is_tree else (self.T + 3, ld),
                                  dtype=self.tdtype, device=dev)
        self.obs = torch.zeros((n, self.obs_dim), dtype=torch.float32, device=dev)
        self.reward = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.ep_return = torch.zeros((n,), dtype=torch.float64, device=dev)
        self.stats = torch.zeros((_lib.SN_STATS_WORDS,), dtype=torch.float64, device=dev)
        # device control words (SN_CTL_*): period counter of the graph-replayed step, ticket, sticky flags
        self.ctl = torch.zeros((_lib.SN_CTL_WORDS,), dtype=torch.int32, device=dev)
        self.period = 0
        self.terminal = True          # no episode loaded yet
        self.track_stats = False
        # distribution trees: a library specialised for this very network (specialize.py), if one has been built,
        # serves the period step and the fused rollouts; everything else stays on the general library
        from . import specialize
        self.fast = specialize.load(spec, dtype) if spec.is_tree else None
        self._step_lib = self.fast if self.fast is not None else self.lib

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    _SRC = {torch.int32: _lib.SN_SRC_I32, torch.int64: _lib.SN_SRC_I64, torch.float32: _lib.SN_SRC_F32,
            torch.float64: _lib.SN_SRC_F64}

    def _pack_rows(self, x, width, dst: torch.Tensor, rows: int, row_stride: int):
        """[n, width] host/device array -> `rows` structure-of-arrays rows of `dst` (sn_pack_rows: transpose, dtype
        conversion and zero padding in one launch; the host->device copy before it is the only other step)."""
        t = torch.as_tensor(x)
        if t.dim() != 2 or t.shape[0] != self.n_envs or t.shape[1] != width:
            raise ValueError(f"expected shape ({self.n_envs}, {width}), got {tuple(t.shape)}")
        if self.dtype_code == _lib.SN_I32 and t.is_floating_point():
            if not bool((t == t.round()).all()):
                raise TypeError("integer simulator (dtype='int32') needs integral quantities")
        if t.dtype not in self._SRC:
            t = t.to(torch.float64 if t.is_floating_point() else torch.int64)
        if t.stride(1) != 1:
            t = t.contiguous()
        t = t.to(device=self.device, non_blocking=True)
        src_stride = t.stride(0) if self.n_envs > 1 else width       # (the stride of a size-1 dimension is arbitrary)
        _lib.check(self.lib.sn_pack_rows(self._SRC[t.dtype], C.c_void_p(t.data_ptr()), self.n_envs, width, src_stride,
                                         self.dtype_code, C.c_void_p(dst.data_ptr()), rows, self.ld, row_stride,
                                         self._stream()))
        return t            # keeps the staging tensor alive until the caller returns (the launch is asynchronous)

    # ------------------------------------------------------------------ episode inputs
    def load_episode(self, init, demand):
        """init [n, init_words] (19: inv[4]; so[k][j]; sh[k][j] packed for the beer game's lead times; 36:
        inv[4]; so[k][0..3]; sh[k][0..3] for other lead times), demand [n, >=T+1] (utils/demands.py streams)."""
        d = torch.as_tensor(demand)
        rows = self.T + 3
        if self.spec.is_tree:            # demand [n, n_sources, >=T+1]: one stream per demand source, node order
            if d.dim() != 3 or d.shape[1] != self.n_sources or d.shape[2] < self.T + 1:
                raise ValueError(f"demand must be [n_envs, {self.n_sources}, >= {self.T + 1}]")
            width = min(d.shape[2], rows)
            keep = [self._pack_rows(init, self.init_words, self.init, self.init_words, self.ld)]
            for s in range(self.n_sources):
                keep.append(self._pack_rows(d[:, s, :width], width, self.demand[0, s], rows, self.n_sources * self.ld))
            return
        if d.dim() != 2 or d.shape[1] < self.T + 1:
            raise ValueError(f"demand must be [n_envs, >= {self.T + 1}]")
        width = min(d.shape[1], rows)
        keep = [self._pack_rows(init, self.init_words, self.init, self.init_words, self.ld),       # noqa: F841
                self._pack_rows(d[:, :width], width, self.demand, rows, self.ld)]

    # ------------------------------------------------------------------ env API
    def reset(self, obs_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """envs.py:64-80 for all envs; returns obs float32 [n, obs_dim] (device)."""
        obs = self.obs if obs_out is None else obs_out
        _lib.check(self.lib.sn_reset(C.byref(self._c_spec), self.dtype_code, self.n_envs, self.ld, _ptr(self.init),
                                     _ptr(self.demand[0]), _ptr(self.state), _ptr(obs), self._stream()))
        self.period = 0
        self.terminal = False
        self.ep_return.zero_()
        if self._graph_period_valid:
            self.ctl[_lib.SN_CTL_PERIOD:_lib.SN_CTL_PERIOD + 1].zero_()
        return obs

    def _check_actions(self, actions: torch.Tensor, lead=()) -> torch.Tensor:
        shape = tuple(lead) + (self.n_envs, self.n_agents)
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dim() == len(shape) - 1 and self.n_agents == 1:
            actions = actions.unsqueeze(-1)
        if tuple(actions.shape) != shape:
            raise ValueError(f"actions must have shape {shape}, got {tuple(actions.shape)}")
        if self.dtype_code == _lib.SN_I32 and actions.is_floating_point():
            if not bool((actions == actions.round()).all()):
                raise TypeError("integer simulator (dtype='int32') needs integral order quantities")
        return actions.to(device=self.device, dtype=self.tdtype).contiguous()

    def step(self, actions, obs_out=None, reward_out=None, reward64_out=None, cost_out=None, env_reward_out=None):
        """envs.py:82-113 for all envs.  Returns (obs, reward, done) with obs/reward on the device;
        `done` is a python bool because all envs run in lockstep.  `env_reward_out` (multi-item batches):
        float32 [n_envs / n_items], the per-env sum of the items' rewards, reduced inside the kernel."""
        if self.terminal:
            raise ValueError("Cannot take action when the state is terminal.")       # envs.py:88-89
        a = self._check_actions(actions)
        obs = self.obs if obs_out is None else obs_out
        rew = self.reward if reward_out is None else reward_out
        io = _lib.SnStepIo()
        io.state, io.actions, io.demand = self.state.data_ptr(), a.data_ptr(), self.demand[self.period + 1].data_ptr()
        io.period = self.period
        io.obs, io.reward = obs.data_ptr(), rew.data_ptr()
        io.env_reward = None if env_reward_out is None else env_reward_out.data_ptr()
        io.reward64 = None if reward64_out is None else reward64_out.data_ptr()
        io.cost = None if cost_out is None else cost_out.data_ptr()
        io.ep_return = self.ep_return.data_ptr()
        io.stats = self.stats.data_ptr() if self.track_stats else None
        _lib.check(self._step_lib.sn_step_ex(C.byref(self._c_spec), self.dtype_code, self.n_envs, self.ld, C.byref(io),
                                             self._stream()), self._step_lib)
        self.period += 1
        self.terminal = self.period >= self.T
        if self._graph_period_valid:                # keep the device counter of the graph-replayed step in sync
            self.ctl[_lib.SN_CTL_PERIOD:_lib.SN_CTL_PERIOD + 1].fill_(self.period)
        return obs, rew, self.terminal

    # ------------------------------------------------------------------ graph-replayed step (device period counter)
    _graph_period_valid = False

    def launch_step_device_period(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor,
                                  env_reward: Optional[torch.Tensor] = None, vecnorm: Optional[dict] = None):
        """One period step whose period is read from (and advanced in) device memory (`self.ctl`, sn_step_ex):
        the launch has no host-side state, so it can be captured ONCE in a CUDA graph and replayed for every
        period of every episode.  All tensors must keep their addresses between replays.  `vecnorm` = dict(scratch,
        obs_rms, ret_rms, returns, gamma) fuses the VecNormalize statistics update into the launch.
        The caller mirrors the period on the host with `advance_host_period()` after each replay and must not
        replay at the terminal period (the launch would raise SN_FLAG_STEP_AFTER_TERMINAL and do nothing)."""
        io = _lib.SnStepIo()
        io.state, io.actions, io.demand = self.state.data_ptr(), actions.data_ptr(), self.demand.data_ptr()
        io.ctl = self.ctl.data_ptr()
        io.obs, io.reward = obs.data_ptr(), reward.data_ptr()
        io.env_reward = None if env_reward is None else env_reward.data_ptr()
        io.ep_return = self.ep_return.data_ptr()
        io.stats = self.stats.data_ptr() if self.track_stats else None
        if vecnorm is not None:
            ptr = lambda k: None if vecnorm.get(k) is None else vecnorm[k].data_ptr()       # noqa: E731
            io.vn_scratch, io.vn_obs_rms, io.vn_ret_rms = ptr("scratch"), ptr("obs_rms"), ptr("ret_rms")
            io.vn_returns, io.vn_gamma = ptr("returns"), float(vecnorm["gamma"])
            io.vn_nobs, io.vn_nrew = ptr("nobs"), ptr("nrew")       # one-kernel mode: normalised outputs of this launch
            if vecnorm.get("nobs") is not None:
                io.vn_clip_obs, io.vn_clip_reward = float(vecnorm["clip_obs"]), float(vecnorm["clip_reward"])
                io.vn_epsilon = float(vecnorm["epsilon"])
        _lib.check(self._step_lib.sn_step_ex(C.byref(self._c_spec), self.dtype_code, self.n_envs, self.ld, C.byref(io),
                                             self._stream()), self._step_lib)

    def _sync_device_period(self):
        """Keeps the device counter of the graph-replayed step equal to the host's period after host-driven calls."""
        if self._graph_period_valid:
            self.ctl[_lib.SN_CTL_PERIOD:_lib.SN_CTL_PERIOD + 1].fill_(self.period)

    def advance_host_period(self):
        """Host mirror of what a replayed `launch_step_device_period` did on the device."""
        self.period += 1
        self.terminal = self.period >= self.T

    def device_flags(self) -> int:
        """Sticky SN_FLAG_* bits raised by device-period launches (synchronises)."""
        return int(self.ctl[_lib.SN_CTL_FLAGS].item())

    def observe(self, obs_out=None) -> torch.Tensor:
        obs = self.obs if obs_out is None else obs_out
        _lib.check(self.lib.sn_observe(C.byref(self._c_spec), self.dtype_code, self.n_envs, self.ld, _ptr(self.state),
                                       _ptr(obs), self._stream()))
        return obs

    def rollout(self, n_steps: Optional[int] = None, actions=None, policy: Optional[BasePolicyParams] = None,
                record_obs=False, record_reward=False, record_actions=False, traj_obs=None, traj_reward=None,
                traj_actions=None):
        """n_steps fused periods.  Exactly one of `actions` ([n_steps, n, n_agents]) / `policy`."""
        if self.terminal:
            raise ValueError("Cannot take action when the state is terminal.")
        if n_steps is None:
            n_steps = self.T - self.period
        if (actions is None) == (policy is None):
            raise ValueError("give exactly one of actions= or policy=")
        n, dev = self.n_envs, self.device
        if actions is not None:
            a = self._check_actions(actions, lead=(n_steps,))
            kind, pol = _lib.SN_POLICY_ACTIONS, None
        else:
            if policy.n_cols != self.n_agents:
                raise ValueError(f"policy has {policy.n_cols} targets, env has {self.n_agents} action columns")
            a, kind, pol = None, _lib.SN_POLICY_BASE_STOCK, C.byref(policy.c)
        if record_obs and traj_obs is None:
            traj_obs = torch.empty((n_steps, n, self.obs_dim), dtype=torch.float32, device=dev)
        if record_reward and traj_reward is None:
            traj_reward = torch.empty((n_steps, n), dtype=torch.float32, device=dev)
        if record_actions and traj_actions is None:
            traj_actions = torch.empty((n_steps, n, self.n_agents), dtype=torch.float32, device=dev)
        _lib.check(self._step_lib.sn_rollout(
            C.byref(self._c_spec), self.dtype_code, n, self.ld, self.period, n_steps, _ptr(self.state),
            _ptr(self.demand), kind, _ptr(a), pol, _ptr(traj_obs), _ptr(traj_reward), _ptr(traj_actions),
            _ptr(self.obs), _ptr(self.ep_return), _ptr(self.stats) if self.track_stats else None, self._stream()),
            self._step_lib)
        self.period += n_steps
        self.terminal = self.period >= self.T
        self._sync_device_period()
        return dict(obs=self.obs, traj_obs=traj_obs, traj_reward=traj_reward, traj_actions=traj_actions,
                    done=self.terminal)

    def episode(self, n_steps: Optional[int] = None, actions=None, policy: Optional[BasePolicyParams] = None,
                record_obs=False, record_reward=False, record_actions=False, traj_obs=None, traj_reward=None,
                traj_actions=None, obs0_out=None, write_state=True):
        """reset() + rollout(n_steps) fused in one launch (sn_episode): the loaded episode is played
        from period 0 with the state kept on-chip.  Returns the same dict as rollout() plus obs0."""
        if n_steps is None:
            n_steps = self.T
        if (actions is None) == (policy is None):
            raise ValueError("give exactly one of actions= or policy=")
        n, dev = self.n_envs, self.device
        if actions is not None:
            a = self._check_actions(actions, lead=(n_steps,))
            kind, pol = _lib.SN_POLICY_ACTIONS, None
        else:
            if policy.n_cols != self.n_agents:
                raise ValueError(f"policy has {policy.n_cols} targets, env has {self.n_agents} action columns")
            a, kind, pol = None, _lib.SN_POLICY_BASE_STOCK, C.byref(policy.c)
        if record_obs and traj_obs is None:
            traj_obs = torch.empty((n_steps, n, self.obs_dim), dtype=torch.float32, device=dev)
        if record_reward and traj_reward is None:
            traj_reward = torch.empty((n_steps, n), dtype=torch.float32, device=dev)
        if record_actions and traj_actions is None:
            traj_actions = torch.empty((n_steps, n, self.n_agents), dtype=torch.float32, device=dev)
        _lib.check(self._step_lib.sn_episode(
            C.byref(self._c_spec), self.dtype_code, n, self.ld, n_steps, _ptr(self.init), _ptr(self.demand), kind,
            _ptr(a), pol, _ptr(obs0_out), _ptr(traj_obs), _ptr(traj_reward), _ptr(traj_actions), _ptr(self.obs),
            _ptr(self.state) if write_state else None, _ptr(self.ep_return),
            _ptr(self.stats) if self.track_stats else None, self._stream()), self._step_lib)
        self.period = n_steps
        self.terminal = self.period >= self.T
        self._sync_device_period()
        return dict(obs=self.obs, obs0=obs0_out, traj_obs=traj_obs, traj_reward=traj_reward,
                    traj_actions=traj_actions, done=self.terminal)

    def episode_host(self, h_init, h_demand, h_actions, h_traj_obs, h_traj_reward, chunks: int = 4):
        """The reference-shaped call with HOST buffers: numpy/pinned-torch inputs in, full trajectory out
        on the host (what `for t: obs, r, d, _ = venv.step(a[t])` hands back), pipelined so that PCIe
        runs in both directions at once: the episode is cut into `chunks` spans of periods; actions of
        span c+1 travel host->device and the trajectory of span c-1 travels device->host while span c
        is computed (sn_episode for the first span, sn_rollout for the others).  Synchronises before
        returning.  h_actions [T, n, n_agents] / h_traj_obs [T, n, obs_dim] f32 / h_traj_reward [T, n] f32
        must be pinned torch tensors for the copies to overlap."""
        T, n, dev = self.T, self.n_envs, self.device
        h_actions = torch.as_tensor(h_actions)
        if tuple(h_actions.shape) != (T, n, self.n_agents):
            raise ValueError(f"h_actions must have shape {(T, n, self.n_agents)}")
        if h_actions.dtype != self.tdtype:
            raise TypeError(f"h_actions must have dtype {self.tdtype} for the pipelined host path")
        if not hasattr(self, "_pipe"):
            self._pipe = dict(h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev),
                              act=torch.empty((T, n, self.n_agents), dtype=self.tdtype, device=dev),
                              obs=torch.empty((T, n, self.obs_dim), dtype=torch.float32, device=dev),
                              rew=torch.empty((T, n), dtype=torch.float32, device=dev))
        pp = self._pipe
        main = torch.cuda.current_stream(dev)
        chunks = max(1, min(int(chunks), T))
        bounds = [(c * T) // chunks for c in range(chunks + 1)]
        pp["h2d"].wait_stream(main)
        ev_in = []
        with torch.cuda.stream(pp["h2d"]):
            for c in range(chunks):
                t0, t1 = bounds[c], bounds[c + 1]
                pp["act"][t0:t1].copy_(h_actions[t0:t1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(pp["h2d"])
                ev_in.append(ev)
        self.load_episode(h_init, h_demand)                  # H2D init + demand on the main stream
        for c in range(chunks):
            t0, t1 = bounds[c], bounds[c + 1]
            main.wait_event(ev_in[c])
            if c == 0:
                self.episode(t1, actions=pp["act"][:t1], traj_obs=pp["obs"][:t1], traj_reward=pp["rew"][:t1])
            else:
                self.rollout(t1 - t0, actions=pp["act"][t0:t1], traj_obs=pp["obs"][t0:t1], traj_reward=pp["rew"][t0:t1])
            ev = torch.cuda.Event()
            ev.record(main)
            pp["d2h"].wait_event(ev)
            with torch.cuda.stream(pp["d2h"]):
                h_traj_obs[t0:t1].copy_(pp["obs"][t0:t1], non_blocking=True)
                h_traj_reward[t0:t1].copy_(pp["rew"][t0:t1], non_blocking=True)
        main.wait_stream(pp["d2h"])
        main.synchronize()
        return h_traj_obs, h_traj_reward

    # ------------------------------------------------------------------ snapshot
    def get_state(self):
        return dict(state=self.state.clone(), demand=self.demand.clone(), init=self.init.clone(),
                    ep_return=self.ep_return.clone(), period=self.period, terminal=self.terminal)

    def set_state(self, snap):
        self.state.copy_(snap["state"]); self.demand.copy_(snap["demand"]); self.init.copy_(snap["init"])
        self.ep_return.copy_(snap["ep_return"])
        self.period, self.terminal = int(snap["period"]), bool(snap["terminal"])
        if self._graph_period_valid:
            self.ctl[_lib.SN_CTL_PERIOD:_lib.SN_CTL_PERIOD + 1].fill_(self.period)


def heuristic_actions(policy: BasePolicyParams, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """BaseStockPolicy.get_order_quantity (utils/heuristics.py:43-57,75) on a device batch."""
    lib = _lib.load()
    if obs.device.type != "cuda":
        raise RuntimeError("heuristic_actions needs a CUDA tensor; there is no CPU fallback")
    if obs.dim() != 2 or obs.shape[1] != 7 * policy.n_cols:
        raise ValueError(f"obs must be [n, {7 * policy.n_cols}], got {tuple(obs.shape)}")
    obs = obs.to(torch.float32).contiguous()
    n = obs.shape[0]
    if out is None:
        out = torch.empty((n, policy.n_cols), dtype=torch.float32, device=obs.device)
    stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
    _lib.check(lib.sn_heuristic(C.byref(policy.c), policy.n_cols, n, _ptr(obs), _ptr(out), stream))
    return out
