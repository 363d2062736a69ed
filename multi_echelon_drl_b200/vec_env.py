"""SupplyNetVecEnv: the stable-baselines3 VecEnv surface on top of the CUDA simulator.

Stands in for `VecNormalize?(make_vec_env(env_factory, n_env))` = DummyVecEnv[Monitor[TimeLimit[env]]]
of the reference's trainers (train_td3.py:82-101, train_a2c.py:68-86, train_dqn.py:69-87) with
thousands to millions of envs instead of 2.  Observations, rewards and dones are torch tensors that
already live on the device (`output="torch"`, north_star) or numpy arrays staged through pinned host
memory (`output="numpy"`, what SB3 itself consumes).

SB3 1.x semantics reproduced: old 4-tuple step API, auto-reset on done with
`infos[i]["terminal_observation"]`, Monitor's `infos[i]["episode"] = {"r", "l", "t"}`,
`infos[i]["TimeLimit.truncated"] = False` (the env terminates itself at max_episode_steps == the
registry's TimeLimit, register_envs.py:77).  All envs share one period counter: episodes have a
fixed length, so a batch that starts together ends together.
"""
from __future__ import annotations

import time
from collections.abc import Sequence
from typing import Optional

import numpy as np
import torch

import ctypes as C

from . import _lib, spaces
from .simulator import BatchedSupplyNetwork
from .spec import ScenarioSpec
from .utils import demands as _dem


class LazyInfos(Sequence):
    """Behaves like SB3's list of per-env info dicts without materialising N dicts per step;
    the batched tensors are available as attributes."""

    def __init__(self, n, terminal_observation=None, episode_return=None, episode_length=0, elapsed=0.0):
        self.n = n
        self.terminal_observation = terminal_observation      # [n, obs_dim] or None
        self.episode_return = episode_return                  # [n] float64 or None
        self.episode_length = episode_length
        self.elapsed = elapsed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        if not -self.n <= i < self.n:
            raise IndexError(i)
        if self.terminal_observation is None:
            return {}
        return {"terminal_observation": self.terminal_observation[i],
                "episode": {"r": round(float(self.episode_return[i]), 6), "l": self.episode_length,
                            "t": round(self.elapsed, 6)},
                "TimeLimit.truncated": False}


try:                                       # pragma: no cover - stable-baselines3 is not installed in this image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase      # type: ignore
except Exception:                          # noqa: BLE001
    _VecEnvBase = object                   # duck type only


class SupplyNetVecEnv(_VecEnvBase):
    """VecEnv duck type; a real `stable_baselines3.common.vec_env.VecEnv` subclass when SB3 is importable, so
    that SB3 algorithms accept it without re-wrapping (use output="numpy" for them).  VecEnv duck type (num_envs, observation_space, action_space, reset, step_async, step_wait, step,
    close, seed, get_attr, set_attr, env_method, env_is_wrapped)."""

    metadata = {"render.modes": []}
    kRing = 3

    def __init__(self, spec: ScenarioSpec, n_envs: int, device="cuda", dtype="float32", seed: int = 0,
                 episode_source: str = "device", output: str = "torch", box_action_space: bool = True,
                 multi_discrete: bool = False, use_graph: bool = False):
        """use_graph (output="torch"): non-terminal steps replay ONE captured CUDA graph (the period lives in device
        memory, sn_step_ex) instead of issuing the launch from Python; a wrapping device VecNormalize joins the same
        graph with its statistics fused into the step kernel.  Actions written into `action_buffer` skip a copy."""
        if episode_source not in ("device", "numpy_bulk", "numpy_seeded"):
            raise ValueError("episode_source must be 'device', 'numpy_bulk' or 'numpy_seeded'")
        if output not in ("torch", "numpy"):
            raise ValueError("output must be 'torch' or 'numpy'")
        if spec.n_agents > 1 and not box_action_space and not multi_discrete:
            raise ValueError("length of agent_managed_facilities >= 1, only box_action_space is allowed. "
                             "Please specify box_action_space=True")                    # envs.py:175-179
        self.spec = spec
        self.num_envs = int(n_envs)
        self.render_mode = None
        self.sim = BatchedSupplyNetwork(spec, n_envs, device, dtype)
        self.device = self.sim.device
        self.output = output
        self.use_graph = bool(use_graph) and output == "torch"
        self._graph, self._graph_key, self._vn = None, None, None
        self._act_buf = torch.zeros((int(n_envs), spec.n_agents), dtype=self.sim.tdtype, device=self.sim.device)
        self.last_step_graphed = False
        self.episode_source = episode_source
        # spaces exactly as the factories build them (envs.py:269-277, 397-405, 587-588)
        if multi_discrete:
            self.action_space = spaces.MultiDiscrete([int(spec.action_high) + 1] * spec.n_agents)
        elif box_action_space:
            self.action_space = spaces.Box(spec.action_low, spec.action_high, shape=(spec.n_agents,))
        else:
            self.action_space = spaces.Discrete(int(spec.action_high) + 1)
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(spec.obs_dim,))
        if _VecEnvBase is not object:      # pragma: no cover - let SB3 set up its own bookkeeping (reset_infos, ...)
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)
        self._actions = None
        self._t0 = time.time()
        self.episodes_done = 0
        n, od = self.num_envs, spec.obs_dim
        self._obs = torch.zeros((n, od), dtype=torch.float32, device=self.device)
        self._term_obs = torch.zeros((n, od), dtype=torch.float32, device=self.device)
        self._all_false = torch.zeros((n,), dtype=torch.bool, device=self.device)
        self._all_true = torch.ones((n,), dtype=torch.bool, device=self.device)
        self._no_infos = LazyInfos(n)
        self._last_returns = torch.zeros((n,), dtype=torch.float64, device=self.device)
        if output == "numpy":
            self._h_act = torch.zeros((n, spec.n_agents), dtype=self.sim.tdtype).pin_memory()
            self._h_act_np = self._h_act.numpy()
            self._d_act = torch.zeros((n, spec.n_agents), dtype=self.sim.tdtype, device=self.device)
            # Results are handed out as views of pinned staging buffers, rotated over kRing sets: an array returned
            # by step()/reset() stays untouched for the next kRing-1 calls (SB3 keeps `_last_obs` across one
            # env.step(); DummyVecEnv allocates fresh arrays instead, which at this batch size would cost a
            # host memcpy of megabytes per step).  Copy what must live longer.
            self._ring = [(torch.zeros((n, od), dtype=torch.float32).pin_memory(),
                           torch.zeros((n,), dtype=torch.float32).pin_memory()) for _ in range(self.kRing)]
            self._ring_pos = 0
            self._h_term = torch.zeros((n, od), dtype=torch.float32).pin_memory()
        self.seed(seed)

    # ------------------------------------------------------------------ episode inputs
    def seed(self, seed: Optional[int] = None):
        """Unlike the reference (whose env.seed only seeds an unused generator, envs.py:115-117, Q9)
        this really seeds the per-env demand / initial-state streams."""
        seed = 0 if seed is None else int(seed)
        self._seed = seed
        if self.episode_source == "device":
            self._ctr = 0                      # Philox counter offset: advances with every generated table
        elif self.episode_source == "numpy_bulk":
            self._rs = np.random.RandomState(seed)
        else:
            self._streams = [np.random.RandomState(seed + i) for i in range(self.num_envs)]
        return [seed + i for i in range(self.num_envs)]

    def _next_episode(self):
        sp, sim = self.spec, self.sim
        if self.episode_source == "device":
            if getattr(self, "_tables_ahead", 0):      # run_policy_episodes left the next episode's tables in place
                self._ctr += self._tables_ahead
                self._tables_ahead = 0
                return
            self._device_episode()
        elif self.episode_source == "numpy_bulk":
            init, demand = _dem.bulk_episode_inputs(sp, self.num_envs, self._rs)
            sim.load_episode(init, demand)
        else:
            init, demand = _dem.seeded_episode_inputs_from_streams(sp, self._streams)
            sim.load_episode(init, demand)

    def _device_episode(self, counter=None, rel0=0, advance=True):
        """Same distributions as utils/demands.py / Arc.reset / Node.reset, drawn on the device with
        the library's Philox-4x32-10 counter streams directly in the structure-of-arrays layout (throughput
        path; not stream-compatible with numpy - use episode_source='numpy_seeded' for reference parity).
        The stream position is `self._ctr` (host) or, with `counter` (device uint64 tensor), *counter + rel0: the
        form a captured graph replays.  Returns the number of counter steps one episode consumes."""
        sp, sim = self.spec, self.sim
        lib, dt, st = sim.lib, sim.dtype_code, sim._stream()
        a, b = sp.demand_params
        seed = self._seed & 0xFFFFFFFFFFFFFFFF
        cptr = None if counter is None else C.c_void_p(counter.data_ptr())
        pos = [rel0 if counter is not None else self._ctr]

        def fill_uniform(t, lo, hi):
            _lib.check(lib.sn_fill_uniform_at(dt, C.c_void_p(t.data_ptr()), t.numel(), int(lo), int(hi), seed, cptr, pos[0], st))
            pos[0] += (t.numel() + 3) // 4

        def fill_normal(t, mean, sd):
            _lib.check(lib.sn_fill_normal_demand_at(dt, C.c_void_p(t.data_ptr()), t.numel(), float(mean), float(sd), seed,
                                                    cptr, pos[0], st))
            pos[0] += (t.numel() + 3) // 4

        # one launch per table, written in place in the structure-of-arrays layout (sn_fill_* in random_fill.cu)
        if sp.is_tree and len(set(sp.source_demand_params)) > 1:
            # demand sources with their own parameters: each source's rows are drawn into a contiguous scratch
            # table from the same seeded Philox counter stream, then copied into its strided slice of the table
            if not hasattr(self, "_src_scratch"):
                self._src_scratch = torch.empty((sim.demand.shape[0], sim.ld), dtype=sim.tdtype, device=self.device)
            t = self._src_scratch
            for s, (pa, pb) in enumerate(sp.source_demand_params):
                if sp.demand_pattern == "uniform":
                    fill_uniform(t, pa, int(pb) + 1)
                elif sp.demand_pattern == "normal":
                    fill_normal(t, pa, pb)
                else:
                    raise NotImplementedError("tree networks draw uniform or normal demand")
                sim.demand[:, s].copy_(t)
        elif sp.demand_pattern == "uniform":
            fill_uniform(sim.demand, a, int(b) + 1)
        elif sp.demand_pattern == "normal":
            fill_normal(sim.demand, a, b)
        else:
            sim.demand[:4].fill_(4)
            sim.demand[4:].fill_(8)
        if sp.random_init:
            n_inv = 8 if sp.n_facilities > 4 else 4          # inventory rows of the init table (utils/demands.init_layout)
            fill_uniform(sim.init[:n_inv], sp.init_inventory[0], sp.init_inventory[1])
            fill_uniform(sim.init[n_inv:], sp.init_pipeline[0], sp.init_pipeline[1])
            # (for non-standard lead times the kernels ignore pipeline slots beyond an arc's lead time)
        else:
            sim.init.copy_(torch.as_tensor(_dem._fixed_init(sp), device=self.device).to(sim.tdtype)[:, None]
                           .expand(sim.init_words, sim.ld))
        total = pos[0] - (rel0 if counter is not None else self._ctr)
        if counter is None and advance:
            self._ctr = pos[0]
        return total

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        self._next_episode()
        self.sim.reset(self._obs)
        self._t0 = time.time()
        return self._out_obs()

    def step_async(self, actions):
        self._actions = actions

    # ------------------------------------------------------------------ graph-replayed step
    @property
    def action_buffer(self):
        """Static [n_envs, n_agents] action tensor of the graph-replayed step (quantity dtype)."""
        return self._act_buf

    def attach_vecnorm(self, vn):
        """A device VecNormalize wrapping this env joins the step graph (vec_normalize.py)."""
        self._vn, self._graph = vn, None

    def _capture(self, key):
        sim, vn = self.sim, self._vn
        sim._graph_period_valid = True
        sim.ctl[0:1].fill_(sim.period)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            sim.launch_step_device_period(self._act_buf, self._obs, sim.reward,
                                          vecnorm=vn.fused_args(key) if (vn is not None and (key[0] or key[2])) else None)
            if vn is not None and key[1] and not key[2]:
                vn.graph_tail()
        self._graph, self._graph_key = g, key

    def _graph_step(self, actions):
        """Non-terminal period through the captured graph.  key = VecNormalize.graph_key()."""
        sim, vn = self.sim, self._vn
        key = vn.graph_key() if vn is not None else (False, False, False)
        if actions is not self._act_buf and not (isinstance(actions, torch.Tensor) and actions.data_ptr() == self._act_buf.data_ptr()):
            self._act_buf.copy_(sim._check_actions(actions))
        if self._graph is None or self._graph_key != key:
            self._capture(key)
        self._graph.replay()
        sim.advance_host_period()
        self.last_step_graphed = True
        return self._obs, sim.reward, self._all_false, self._no_infos

    def step_wait(self):
        sim = self.sim
        actions = self._actions
        self.last_step_graphed = False
        if self.use_graph and not sim.terminal and sim.period + 1 < sim.T:
            return self._graph_step(actions)
        if self.output == "numpy":
            a = np.asarray(actions)
            if a.ndim == 1:
                a = a[:, None]
            if a.shape != self._h_act_np.shape:
                raise ValueError(f"actions must have shape {self._h_act_np.shape}, got {a.shape}")
            if sim.dtype_code == _lib.SN_I32 and a.dtype.kind == "f" and not np.array_equal(a, np.round(a)):
                raise TypeError("integer simulator (dtype='int32') needs integral order quantities")
            # one pass, dtype conversion included, straight into pinned memory (torch's copy is 2-3x numpy's copyto)
            self._h_act.copy_(torch.from_numpy(a if a.flags.c_contiguous else np.ascontiguousarray(a)))
            self._d_act.copy_(self._h_act, non_blocking=True)
            actions = self._d_act
        elif isinstance(actions, torch.Tensor) and actions.dim() == 1:
            actions = actions[:, None]
        _, rew, done = sim.step(actions, obs_out=self._obs)
        if done:
            self._term_obs.copy_(self._obs)
            self._last_returns.copy_(sim.ep_return)
            self.episodes_done += self.num_envs
            infos = LazyInfos(self.num_envs, self._term_obs, self._last_returns, sim.T, time.time() - self._t0)
            self._next_episode()                    # auto-reset (DummyVecEnv.step_wait)
            sim.reset(self._obs)
            self._t0 = time.time()
        else:
            infos = self._no_infos
        if self.output == "numpy":
            h_obs, h_rew = self._next_ring()
            h_obs.copy_(self._obs, non_blocking=True)
            h_rew.copy_(rew, non_blocking=True)
            if done:
                self._h_term.copy_(self._term_obs, non_blocking=True)
                infos.episode_return = self._last_returns.cpu().numpy()
            torch.cuda.current_stream(self.device).synchronize()
            if done:
                infos.terminal_observation = self._h_term.numpy().copy()
            return (h_obs.numpy(), h_rew.numpy(),
                    np.ones((self.num_envs,), dtype=bool) if done else np.zeros((self.num_envs,), dtype=bool), infos)
        return self._obs, rew, (self._all_true if done else self._all_false), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _out_obs(self):
        if self.output == "numpy":
            h_obs, _ = self._next_ring()
            h_obs.copy_(self._obs, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return h_obs.numpy()
        return self._obs

    def _next_ring(self):
        self._ring_pos = (self._ring_pos + 1) % self.kRing
        return self._ring[self._ring_pos]

    def run_policy_episodes(self, policy, n_episodes: int = 1, use_graph: bool = True):
        """Config-3 mode: play `n_episodes` whole episodes with a base-stock policy evaluated inside the
        rollout kernel (one `sn_episode` launch per episode, state on-chip), drawing fresh demand /
        initial states per episode from this env's episode source.  `policy` is a
        `utils.heuristics.BaseStockPolicy` or `simulator.BasePolicyParams`.  Returns the statistics
        vector (float64 [16], see distributed.py) accumulated over the episodes, on the device.
        `use_graph` (device episode source): pairs of episodes are replayed as ONE captured CUDA graph in which the
        tables of the next episode are drawn (second set of buffers, parallel branch) while the current one is played -
        the rollout is ALU-bound, the fills HBM-write-bound - and the Philox stream position lives in device memory
        (`sn_fill_*_at`, `sn_counter_add`).  Same counters in the same order, hence the same episodes as the loop."""
        params = getattr(policy, "params", policy)
        sim = self.sim
        keep = sim.track_stats
        sim.track_stats = True
        sim.stats.zero_()
        n_episodes = int(n_episodes)
        done = 0
        if use_graph and self.episode_source == "device" and n_episodes >= 2:
            done = self._replay_episode_pairs(params, n_episodes // 2)
        for _ in range(n_episodes - done):
            self._next_episode()
            sim.episode(policy=params, write_state=False)          # statistics incl. episode moments accumulate in-kernel
            self.episodes_done += self.num_envs
        sim.track_stats = keep
        return sim.stats

    def _replay_episode_pairs(self, params, pairs: int) -> int:
        sim, dev = self.sim, self.device
        if not getattr(self, "_tables_ahead", 0):                  # tables of the next episode, position not yet advanced
            self._tables_ahead = self._device_episode(advance=False)
        per = self._tables_ahead
        key = (bytes(params.c), sim.init.data_ptr(), sim.demand.data_ptr(), sim.track_stats)
        if getattr(self, "_pair_key", None) != key:
            if not hasattr(self, "_d_ctr"):
                self._d_ctr = torch.zeros((1,), dtype=torch.int64, device=dev)       # read as uint64 by the fills
                self._alt_tables = (torch.empty_like(sim.init), torch.empty_like(sim.demand))
                self._side = torch.cuda.Stream(dev)
            cur, alt, side = (sim.init, sim.demand), self._alt_tables, self._side
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                main = torch.cuda.current_stream(dev)
                for (play, fill, rel) in ((cur, alt, per), (alt, cur, 2 * per)):
                    side.wait_stream(main)
                    sim.init, sim.demand = fill
                    with torch.cuda.stream(side):
                        self._device_episode(counter=self._d_ctr, rel0=rel)        # next episode's tables, parallel branch
                    sim.init, sim.demand = play
                    sim.episode(policy=params, write_state=False)
                    main.wait_stream(side)
                sim.init, sim.demand = cur
                _lib.check(sim.lib.sn_counter_add(C.c_void_p(self._d_ctr.data_ptr()), 2 * per, sim._stream()))
            self._pair_graph, self._pair_key = g, key
        self._d_ctr.fill_(self._ctr)
        for _ in range(pairs):
            self._pair_graph.replay()
        self._ctr += 2 * per * pairs                                # `cur` again holds the next episode's tables
        self.episodes_done += 2 * pairs * self.num_envs
        return 2 * pairs

    def close(self):
        pass

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [False] * n

    def get_original_obs(self):
        return self._obs

    # convenience mirrors of the single-env attributes consumers read (envs.py:53-60)
    @property
    def period(self):
        return self.sim.period

    @property
    def terminal(self):
        return self.sim.terminal

    @property
    def max_episode_steps(self):
        return self.spec.max_episode_steps
