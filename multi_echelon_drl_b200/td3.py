"""TD3 + heuristic-guided exploration on the device VecEnv (BASELINE config 5).

Plain-PyTorch counterpart of the reference's training stack for the continuous-action case:
`HgeTD3` (hge.py:47-172, a stable-baselines3 TD3 subclass), `HgeRateCallback`
(utils/callbacks.py:5-23) and the wiring of train_td3.py:100-143.  SB3 is third-party and not
installed here; behaviour follows SB3 1.x TD3 as documented: twin critics,
delayed policy updates (policy_delay 2), target policy smoothing (noise 0.2, clip 0.5), Polyak tau,
`learning_starts` uniform-random warm-up, Gaussian action noise added in the scaled [-1,1] action
space (also to heuristic actions, hge.py:129-137), replay buffer holding ORIGINAL observations and
rewards that are normalised with the current VecNormalize statistics at sample time.

Everything stays on the GPU: observations / rewards come from the CUDA env kernels, the heuristic is
`sn_heuristic`, the replay buffer is a set of device tensors, the MLPs are torch.nn modules (the small
actor/critic MLPs stay in PyTorch per north_star).  Parity with SB3's numerics is unpinned.

Difference on purpose: SB3 runs `gradient_steps = train_freq * n_envs` updates per rollout
(gradient_steps=-1, hge.py:66); with thousands of envs that is impractical, so `gradient_steps` is
explicit here (default: train_freq).  `gradient_steps=-1` reproduces SB3's rule.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .hge import HgeRateSchedule, hge_predict
from .vec_normalize import VecNormalize


def mlp(inp, hidden, out, out_act=None):
    layers, d = [], inp
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ReLU()]
        d = h
    layers.append(nn.Linear(d, out))
    if out_act is not None:
        layers.append(out_act)
    return nn.Sequential(*layers)


class TwinQ(nn.Module):
    """The twin critics (SB3 `ContinuousCritic(n_critics=2)`): two MLPs of the same shape, initialised like nn.Linear,
    with their parameters STACKED so that both are evaluated by one batched GEMM per layer - the update graph is bound
    by its number of small launches, and this halves the critics' share.  `both(x)` -> [2, B]; `self[i](x)` -> [B, 1]
    is critic i alone (the actor loss uses critic 0 only, SB3 `q1_forward`)."""

    def __init__(self, d_in, hidden, n=2):
        super().__init__()
        dims = [d_in, *hidden, 1]
        self.n = n
        self.weights, self.biases = nn.ParameterList(), nn.ParameterList()
        for a, b in zip(dims[:-1], dims[1:]):
            bound = 1.0 / math.sqrt(a)                  # nn.Linear's default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
            self.weights.append(nn.Parameter(torch.empty(n, a, b).uniform_(-bound, bound)))
            self.biases.append(nn.Parameter(torch.empty(n, 1, b).uniform_(-bound, bound)))

    def both(self, x):
        h = x.unsqueeze(0).expand(self.n, -1, -1)
        last = len(self.weights) - 1
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = torch.baddbmm(b, h, w)
            if i < last:
                h = torch.relu(h)
        return h.squeeze(-1)

    def one(self, i, x):
        last = len(self.weights) - 1
        for j, (w, b) in enumerate(zip(self.weights, self.biases)):
            x = torch.addmm(b[i], x, w[i])
            if j < last:
                x = torch.relu(x)
        return x

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return lambda x: self.one(i, x)


class ReplayBuffer:
    """Ring buffer of device tensors; `add` takes one vector step (n_envs transitions)."""

    def __init__(self, capacity, obs_dim, act_dim, device):
        self.capacity, self.pos, self.full = int(capacity), 0, False
        kw = dict(dtype=torch.float32, device=device)
        self.obs = torch.zeros((capacity, obs_dim), **kw)
        self.next_obs = torch.zeros((capacity, obs_dim), **kw)
        self.act = torch.zeros((capacity, act_dim), **kw)
        self.rew = torch.zeros((capacity,), **kw)
        self.done = torch.zeros((capacity,), **kw)
        self.size_t = torch.zeros((), dtype=torch.float64, device=device)   # len(self) on the device (graph replays read it)
        self.pos_t = torch.zeros((), dtype=torch.int64, device=device)      # write position on the device (add_on_device)
        self.draws_t = torch.zeros((), dtype=torch.int64, device=device)    # stream position of the device sampler
        self.sample_seed = 0
        self._c, self._lib = None, None

    def __len__(self):
        return self.capacity if self.full else self.pos

    def add(self, obs, next_obs, act, rew, done):
        n = obs.shape[0]
        if n > self.capacity:
            raise ValueError("replay buffer smaller than one vector step")
        idx = (torch.arange(n, device=obs.device) + self.pos) % self.capacity
        self.obs[idx], self.next_obs[idx], self.act[idx] = obs, next_obs, act
        self.rew[idx], self.done[idx] = rew, done.to(torch.float32)
        self.full = self.full or (self.pos + n >= self.capacity)
        self.pos = (self.pos + n) % self.capacity
        self.size_t.fill_(float(len(self)))
        self.pos_t.fill_(self.pos)

    def _c_buffer(self):
        """sn_replay view of the tensors (include/supplynet.h); None off the GPU (unit tests of the host logic)."""
        if self.obs.device.type != "cuda":
            return None
        if self._c is None:
            from . import _lib
            self._lib = _lib.load()
            c = _lib.SnReplay()
            c.obs, c.next_obs, c.act = self.obs.data_ptr(), self.next_obs.data_ptr(), self.act.data_ptr()
            c.rew, c.done = self.rew.data_ptr(), self.done.data_ptr()
            c.capacity, c.obs_dim, c.act_dim = self.capacity, self.obs.shape[1], self.act.shape[1]
            c.pos, c.size = self.pos_t.data_ptr(), self.size_t.data_ptr()
            self._c = c
        return self._c

    @staticmethod
    def _f32(t):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("the device replay buffer takes contiguous float32 tensors")
        return t.data_ptr()

    def add_on_device(self, obs, next_obs, act, rew, done=None, carry=None):
        """`add` with the write position and the length read from / advanced in device memory (sn_replay_add: two
        launches): nothing of it depends on host state, so one captured graph stores every vector step.  `done` None =
        zeros; `carry` (may be `obs` itself) receives next_obs - the loop's current observation for the next step.
        The host mirrors move with `advance_host`."""
        from . import _lib
        c = self._c_buffer()
        n = obs.shape[0]
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        _lib.check(self._lib.sn_replay_add(C.byref(c), n, self._f32(obs), self._f32(next_obs), self._f32(act), self._f32(rew),
                                           None if done is None else self._f32(done),
                                           None if carry is None else self._f32(carry), stream))

    def advance_host(self, n):
        self.full = self.full or (self.pos + n >= self.capacity)
        self.pos = (self.pos + n) % self.capacity

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, len(self), (batch_size,), device=self.obs.device, generator=generator)
        return self.obs[idx], self.act[idx], self.next_obs[idx], self.rew[idx], self.done[idx]

    def sample_on_device(self, batch_size):
        """The same draw with the buffer length read from device memory and the sampler's stream position kept there
        (sn_replay_sample: Philox stream (sample_seed, draws_t), two launches instead of eight indexing kernels): safe
        to capture in a CUDA graph (the length baked into `sample` would go stale across replays)."""
        from . import _lib
        c = self._c_buffer()
        kw = dict(dtype=torch.float32, device=self.obs.device)
        obs, nobs = torch.empty((batch_size, self.obs.shape[1]), **kw), torch.empty((batch_size, self.obs.shape[1]), **kw)
        act, rew, done = torch.empty((batch_size, self.act.shape[1]), **kw), torch.empty(batch_size, **kw), torch.empty(batch_size, **kw)
        stream = C.c_void_p(torch.cuda.current_stream(self.obs.device).cuda_stream)
        _lib.check(self._lib.sn_replay_sample(C.byref(c), batch_size, self.sample_seed, self.draws_t.data_ptr(), obs.data_ptr(),
                                              nobs.data_ptr(), act.data_ptr(), rew.data_ptr(), done.data_ptr(), stream))
        return obs, act, nobs, rew, done

@dataclass
class TD3Config:
    """Defaults = setup_exmaple_centralized.yaml:6-15 through train_td3.py:103-123."""
    batch_size: int = 128
    net_arch: tuple = (768,)              # [width * num_layers]: ONE hidden layer (train_td3.py:103)
    learning_rate: float = 1e-3
    tau: float = 0.01
    train_freq: int = 32
    gradient_steps: int = 32
    action_noise_std: float = 0.4
    gamma: float = 0.95
    hge_rate_at_start: float = 0.5
    hge_decay_periods: int = 1000
    buffer_size: int = 1_000_000
    learning_starts: int = 100
    policy_delay: int = 2
    target_policy_noise: float = 0.2
    target_noise_clip: float = 0.5
    seed: int = 0
    # replay `policy_delay` consecutive updates (sampling, normalisation, both losses, Adam, Polyak) as ONE CUDA graph:
    # the update is ~150 small launches on [128 x 768] tensors, i.e. launch-bound when run eagerly
    cuda_graph: bool = True
    fused_adam: bool = True               # with cuda_graph: torch's single-launch Adam instead of the foreach form


@dataclass
class Timers:
    env_step: float = 0.0
    heuristic: float = 0.0
    policy_forward: float = 0.0
    update: float = 0.0
    events: list = field(default_factory=list)

    def span(self, name):
        return _Span(self, name)

    def flush(self):
        torch.cuda.synchronize()
        for name, a, b in self.events:
            setattr(self, name, getattr(self, name) + a.elapsed_time(b) * 1e-3)
        self.events.clear()


class _Span:
    def __init__(self, timers, name):
        self.t, self.name = timers, name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        b = torch.cuda.Event(enable_timing=True)
        b.record()
        self.t.events.append((self.name, self.a, b))


class HgeTD3:
    """model = HgeTD3(env, heuristic, cfg); model.learn(total_timesteps)."""

    def __init__(self, env, heuristic=None, cfg: Optional[TD3Config] = None, ddp: bool = False):
        self.env, self.heuristic, self.cfg = env, heuristic, cfg or TD3Config()
        c = self.cfg
        self.device = env.device
        self.vec_normalize = env if isinstance(env, VecNormalize) else None
        od, ad = env.observation_space.shape[0], env.action_space.shape[0]
        self.low = torch.as_tensor(np.asarray(env.action_space.low), dtype=torch.float32, device=self.device)
        self.high = torch.as_tensor(np.asarray(env.action_space.high), dtype=torch.float32, device=self.device)
        torch.manual_seed(c.seed)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(c.seed)
        self.rng = np.random.RandomState(c.seed)
        mk = lambda m: m.to(self.device)
        self.actor = mk(mlp(od, c.net_arch, ad, nn.Tanh()))
        self.actor_target = mk(mlp(od, c.net_arch, ad, nn.Tanh()))
        self.critics = mk(TwinQ(od + ad, c.net_arch))
        self.critics_target = mk(TwinQ(od + ad, c.net_arch))
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critics_target.load_state_dict(self.critics.state_dict())
        if ddp:
            self._broadcast_parameters()
        # fused: one multi-tensor launch per optimizer step instead of a dozen foreach launches (the update graph is
        # bound by the latency of its chain of small kernels)
        adam = dict(lr=c.learning_rate, capturable=True, fused=c.fused_adam) if c.cuda_graph else dict(lr=c.learning_rate)
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), **adam)
        self.critic_opt = torch.optim.Adam(self.critics.parameters(), **adam)
        self._graph, self._graph_warm = None, 0
        self._pi_graphs, self._pi_seen = {}, {}
        self._collect, self._collect_warm, self._orig = None, 0, None
        self.buffer = ReplayBuffer(c.buffer_size, od, ad, self.device)
        self.buffer.sample_seed = int(c.seed)
        self.hge_schedule = HgeRateSchedule(c.hge_rate_at_start, c.hge_decay_periods)
        self.hge_rate = c.hge_rate_at_start if heuristic is not None else 0.0
        self.num_timesteps, self.n_updates = 0, 0
        self.timers = Timers()
        self.episode_returns = []

    # ---- action scaling like SB3's policy.scale_action / unscale_action
    def scale_action(self, a):
        return 2.0 * (a - self.low) / (self.high - self.low) - 1.0

    def unscale_action(self, s):
        return self.low + 0.5 * (s + 1.0) * (self.high - self.low)

    @torch.no_grad()
    def predict(self, observation, deterministic=False, original_obs=None):
        """HgeTD3.predict (hge.py:144-172): one coin per call with probability hge_rate**2."""
        def policy(o):
            with self.timers.span("policy_forward"):
                return self._actor_forward(o)

        if self.heuristic is None or deterministic:
            return policy(observation)

        class _H:
            def get_order_quantity(_, o):
                with self.timers.span("heuristic"):
                    return self.heuristic.get_order_quantity(o)

        a, _ = hge_predict(policy, _H(), observation, original_obs if original_obs is not None else observation,
                           self.hge_rate, deterministic, rng=self.rng)
        return a

    @torch.no_grad()
    def _actor_forward(self, o):
        """unscale(actor(o)).  With cuda_graph the forward over the env's observation buffer (the same device
        address at every step) is replayed as a graph: a handful of tiny GEMM / pointwise launches otherwise
        cost more in launch overhead than in work.  The result is a buffer reused by the next call."""
        if not self.cfg.cuda_graph:
            return self.unscale_action(self.actor(o))
        key = (o.data_ptr(), tuple(o.shape), o.dtype)
        g = self._pi_graphs.get(key)
        if g is None:
            n_seen = self._pi_seen.get(key, 0)
            self._pi_seen[key] = n_seen + 1
            if n_seen < 3 or len(self._pi_graphs) >= 4:      # warm up eagerly first; few distinct buffers only
                return self.unscale_action(self.actor(o))
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.unscale_action(self.actor(o))
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self.unscale_action(self.actor(o))
            g = self._pi_graphs[key] = (graph, out, o)       # keeps `o` alive: its address is baked into the graph
        g[0].replay()
        return g[1]

    @torch.no_grad()
    def _sample_action(self, obs, original_obs):
        c, n = self.cfg, obs.shape[0]
        if self.num_timesteps < c.learning_starts:
            u = torch.rand((n, self.low.shape[0]), device=self.device, generator=self.gen)
            unscaled = self.low + u * (self.high - self.low)
        else:
            unscaled = self.predict(obs, False, original_obs)
        scaled = self.scale_action(unscaled.to(torch.float32))
        if c.action_noise_std > 0:
            noise = torch.randn(scaled.shape, device=self.device, generator=self.gen) * c.action_noise_std
            scaled = torch.clamp(scaled + noise, -1.0, 1.0)
        # the env's static action buffer (graph-replayed step) receives the result directly: no copy launch
        buf = getattr(self.env, "action_buffer", None) if getattr(self.env, "use_graph", False) else None
        if buf is not None and buf.dtype == torch.float32 and buf.shape == scaled.shape:
            torch.add((scaled + 1.0) * (0.5 * (self.high - self.low)), self.low, out=buf)
            return buf, scaled
        return self.unscale_action(scaled), scaled

    # ---- collection as three graph replays per vector step: action (actor or heuristic + noise), env step, store
    def _collect_graph_ok(self, obs, orig):
        """The env hands out fixed device buffers and steps through its own replayed graph (VecNormalize over a
        use_graph SupplyNetVecEnv), exploration is past the uniform warm-up, and a few eager steps have run."""
        c, env = self.cfg, self.vec_normalize
        if not c.cuda_graph or env is None or self.num_timesteps < c.learning_starts:
            return False
        venv = env.venv
        buf = getattr(venv, "action_buffer", None) if getattr(venv, "use_graph", False) else None
        if buf is None or getattr(venv, "_graph", None) is None or buf.dtype != torch.float32 or buf.shape[1] != self.low.shape[0]:
            return False
        cg = self._collect
        if cg is not None:
            return obs is cg.obs and orig is cg.orig and buf is cg.buf
        self._collect_warm += 1
        return self._collect_warm > 3

    def _capture(self, fn):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            fn()
        return graph

    @torch.no_grad()
    def _act_graphed(self, obs, orig):
        """`_sample_action` as one replay: unscale(actor(obs)) or the heuristic's orders, scaled to [-1, 1], Gaussian
        noise (default CUDA generator: graph-safe), clipped; the scaled action stays in `cg.scaled` for the replay
        buffer and the env's action buffer receives the un-scaled one.  The coin is HgeTD3.predict's (hge.py:144-172)."""
        from types import SimpleNamespace
        from .utils.heuristics import BaseStockPolicy
        c, cg = self.cfg, self._collect
        if cg is None:
            buf = self.env.venv.action_buffer
            cg = self._collect = SimpleNamespace(obs=obs, orig=orig, buf=buf, scaled=torch.empty_like(buf), store=None, store_key=None)
            # constants of the two affine maps, kept alive: the graphs read them at every replay
            cg.half_span = 0.5 * (self.high - self.low)
            cg.to_scaled_mul = 2.0 / (self.high - self.low)                   # scale_action(a) = a * mul + add
            cg.to_scaled_add = -2.0 * self.low / (self.high - self.low) - 1.0
            cg.to_order_add = self.low + cg.half_span                          # unscale_action(s) = s * half_span + add

            def explore(scaled):
                """noise, clip, both forms of the action: four launches."""
                if c.action_noise_std > 0:
                    scaled = torch.add(scaled, torch.randn_like(scaled), alpha=c.action_noise_std)
                torch.clamp(scaled, -1.0, 1.0, out=cg.scaled)
                torch.addcmul(cg.to_order_add, cg.scaled, cg.half_span, out=buf)

            def from_orders(orders):
                explore(torch.addcmul(cg.to_scaled_add, orders.to(torch.float32), cg.to_scaled_mul))
            cg.from_orders = from_orders
            # the actor's tanh output IS the scaled action (eager: scale_action(unscale_action(.)), equal up to rounding)
            cg.pi = self._capture(lambda: explore(self.actor(obs)))
            cg.heur = (self._capture(lambda: from_orders(self.heuristic.get_order_quantity(orig)))
                       if isinstance(self.heuristic, BaseStockPolicy) else None)
        p = float(self.hge_rate) ** 2
        if self.heuristic is not None and p > 0.0 and self.rng.rand() < p:
            with self.timers.span("heuristic"):
                if cg.heur is not None:
                    cg.heur.replay()
                else:
                    cg.from_orders(torch.as_tensor(self.heuristic.get_order_quantity(orig), device=self.device))
        else:
            with self.timers.span("policy_forward"):
                cg.pi.replay()
        return cg.buf, cg.scaled

    @torch.no_grad()
    def _store_graphed(self, orig, new_orig, raw_rew, dones):
        """buffer.add(orig, new_orig, scaled action, reward, dones) and orig <- new_orig as one replay (non-terminal
        steps of the env's replayed graph: every operand is a fixed device buffer)."""
        cg = self._collect
        key = (new_orig.data_ptr(), raw_rew.data_ptr())
        if cg.store is None:
            # eagerly once, then the same two launches as a graph: (orig, new_orig, scaled action, reward, done = 0) into
            # the buffer at its device-side position, and orig <- new_orig in the same kernel
            self.buffer.add_on_device(orig, new_orig, cg.scaled, raw_rew, None, carry=orig)
            cg.keep = (new_orig, raw_rew)
            cg.store = self._capture(lambda: self.buffer.add_on_device(orig, new_orig, cg.scaled, raw_rew, None, carry=orig))
            cg.store_key = key
        elif key == cg.store_key:
            cg.store.replay()
        else:
            self.buffer.add_on_device(orig, new_orig, cg.scaled, raw_rew, None, carry=orig)
        self.buffer.advance_host(orig.shape[0])

    def _normalize(self, obs, rew, stats=None):
        if self.vec_normalize is None:
            return obs, rew
        return self.vec_normalize.normalize_obs(obs, stats), self.vec_normalize.normalize_reward(rew, stats)

    def _update(self, do_actor: bool, on_device_sampling: bool = False, stats=None):
        """One TD3 gradient step (SB3 TD3.train body): critics always, actor + Polyak when `do_actor`."""
        c = self.cfg
        gen = None if on_device_sampling else self.gen
        obs, act, nobs, rew, done = (self.buffer.sample_on_device(c.batch_size) if on_device_sampling
                                     else self.buffer.sample(c.batch_size, self.gen))
        obs, rew = self._normalize(obs, rew, stats)
        if self.vec_normalize is not None:
            nobs = self.vec_normalize.normalize_obs(nobs, stats)
        with torch.no_grad():
            # (few launches on purpose: the update graph is bound by the latency of its chain of small kernels)
            noise = torch.empty_like(act).normal_(0.0, c.target_policy_noise, generator=gen
                                                  ).clamp_(-c.target_noise_clip, c.target_noise_clip)
            na = (self.actor_target(nobs) + noise).clamp_(-1.0, 1.0)
            x = torch.cat([nobs, na], dim=1)
            q_next = self.critics_target.both(x).amin(0)
            target = torch.addcmul(rew, torch.rsub(done, 1.0), q_next, value=c.gamma)      # rew + gamma (1 - done) q
        x = torch.cat([obs, act], dim=1)
        qs = self.critics.both(x)                            # sum over the critics of each one's mean squared error
        critic_loss = F.mse_loss(qs, target.unsqueeze(0).expand_as(qs), reduction="sum") / qs.shape[1]
        self.critic_opt.zero_grad(set_to_none=True)
        critic_loss.backward()
        self._allreduce_grads(self.critics)
        self.critic_opt.step()
        if do_actor:
            actor_loss = -self.critics[0](torch.cat([obs, self.actor(obs)], dim=1)).mean()
            self.actor_opt.zero_grad(set_to_none=True)
            actor_loss.backward()
            self._allreduce_grads(self.actor)
            self.actor_opt.step()
            with torch.no_grad():
                ps = list(self.actor.parameters()) + list(self.critics.parameters())
                pts = list(self.actor_target.parameters()) + list(self.critics_target.parameters())
                torch._foreach_mul_(pts, 1 - c.tau)
                torch._foreach_add_(pts, ps, alpha=c.tau)

    def _cycle(self, stats):
        """`policy_delay` updates, the last one with the actor step: the unit that is captured and replayed."""
        for i in range(self.cfg.policy_delay):
            self._update(i == self.cfg.policy_delay - 1, on_device_sampling=True, stats=stats)

    def train(self, gradient_steps):
        c = self.cfg
        cycles = 0
        if c.cuda_graph:
            while gradient_steps and self.n_updates % c.policy_delay:      # align to the actor-update cadence
                self.n_updates += 1
                gradient_steps -= 1
                self._update(self.n_updates % c.policy_delay == 0)
            cycles = gradient_steps // c.policy_delay
        if cycles:
            stats = self.vec_normalize.frozen_stats() if self.vec_normalize is not None else None
            if self._graph is None and self._graph_warm < 3:
                # the first cycles run eagerly on a side stream (optimizer state, cuBLAS workspaces), as
                # torch.cuda.graph requires before capture; they are real updates
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    while cycles and self._graph_warm < 3:
                        self._cycle(stats)
                        self._graph_warm += 1
                        cycles -= 1
                        self.n_updates += c.policy_delay
                        gradient_steps -= c.policy_delay
                torch.cuda.current_stream(self.device).wait_stream(side)
            if cycles and self._graph is None:
                self.critic_opt.zero_grad(set_to_none=True)
                self.actor_opt.zero_grad(set_to_none=True)
                self._graph = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread may query events while this thread captures
                with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                    self._cycle(stats)
            for _ in range(cycles):
                self._graph.replay()
            self.n_updates += cycles * c.policy_delay
            gradient_steps -= cycles * c.policy_delay
        for _ in range(gradient_steps):               # eager path (cuda_graph off, or a remainder)
            self.n_updates += 1
            self._update(self.n_updates % c.policy_delay == 0)

    def _broadcast_parameters(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for m in (self.actor, self.actor_target, self.critics, self.critics_target):
                for p in m.parameters():
                    dist.broadcast(p.data, src=0)

    @staticmethod
    def _allreduce_grads(module):
        """Data-parallel training across GPUs: average gradients over ranks (NCCL over NVLink): one flat buffer, one
        all-reduce that averages in the collective itself, one multi-tensor copy back: three launches per call inside
        the update graph (a division and one copy per parameter tensor before)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        if dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:                                   # gloo has no averaging reduction
            dist.all_reduce(flat)
            flat /= dist.get_world_size()
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)

    def learn(self, total_timesteps, log_interval=None):
        """OffPolicyAlgorithm.learn: collect train_freq vector steps, then train (train_td3.py:138-143)."""
        c, env = self.cfg, self.env
        n = env.num_envs
        obs = env.reset()
        first = env.get_original_obs()
        if self._orig is None or self._orig.shape != first.shape or self._orig.dtype != first.dtype:
            self._orig = first.clone()        # ONE buffer for the model's lifetime: the collection graphs read it
        else:
            self._orig.copy_(first)
        orig = self._orig
        t0 = time.time()
        while self.num_timesteps < total_timesteps:
            for _ in range(c.train_freq):
                self.hge_rate = self.hge_schedule(self.num_timesteps) if self.heuristic is not None else 0.0
                graphed = self._collect_graph_ok(obs, orig)
                action, buffer_action = self._act_graphed(obs, orig) if graphed else self._sample_action(obs, orig)
                with self.timers.span("env_step"):
                    new_obs, rew, dones, infos = env.step(action)
                new_orig = env.get_original_obs()
                raw_rew = env.get_original_reward() if self.vec_normalize is not None else rew
                done = bool(infos.terminal_observation is not None)
                if graphed and not done and new_obs is obs and getattr(env.venv, "last_step_graphed", False):
                    self._store_graphed(orig, new_orig, raw_rew, dones)
                    self.num_timesteps += n
                    continue
                # at an episode end the stored next observation is the terminal one, not the reset one
                next_for_buffer = new_orig
                if done:
                    next_for_buffer = infos.terminal_observation if self.vec_normalize is None else \
                        self.vec_normalize.venv._term_obs
                    self.episode_returns.append(float(infos.episode_return.mean().item()))
                self.buffer.add(orig, next_for_buffer, buffer_action, raw_rew, dones)
                obs = new_obs
                orig.copy_(new_orig)          # `orig` is one fixed buffer (the collection graphs read it)
                self.num_timesteps += n
            if self.num_timesteps >= c.learning_starts and len(self.buffer) >= c.batch_size:
                steps = c.gradient_steps if c.gradient_steps >= 0 else c.train_freq * n
                with self.timers.span("update"):
                    self.train(steps)
            if log_interval and (self.num_timesteps // n) % log_interval < c.train_freq:
                self.timers.flush()
                last = self.episode_returns[-1] if self.episode_returns else float("nan")
                print(f"timesteps {self.num_timesteps:>12,}  updates {self.n_updates:>7,}  hge_rate {self.hge_rate:.3f}  "
                      f"mean episode return {last:10.1f}  fps {self.num_timesteps / (time.time() - t0):,.0f}", flush=True)
        self.timers.flush()
        return self

    @torch.no_grad()
    def evaluate(self, eval_env, n_episodes: int = 1):
        """EvalCallback's evaluate_policy (train_td3.py:126-135: 100 deterministic episodes): plays
        `n_episodes` episodes of every env of `eval_env` with the deterministic actor.  `eval_env` is a
        VecNormalize whose observation statistics are synchronised from the training env first (SB3's
        sync_envs_normalization) and frozen; returns are the ORIGINAL (un-normalised) episode returns.
        Returns (mean, std, number of episodes)."""
        if self.vec_normalize is not None and isinstance(eval_env, VecNormalize):
            eval_env._obs_rms.copy_(self.vec_normalize._obs_rms)
            eval_env._ret_rms.copy_(self.vec_normalize._ret_rms)
            eval_env.training = False
        rets = []
        obs = eval_env.reset()
        while len(rets) < n_episodes:
            obs, _, dones, infos = eval_env.step(self.unscale_action(self.actor(obs)))
            if infos.terminal_observation is not None:
                rets.append(infos.episode_return.clone())
        r = torch.cat(rets)
        return float(r.mean().item()), float(r.std().item()), int(r.numel())

    def save(self, path: str):
        """Best-model / env-statistics checkpoint (EvalCallback best_model_save_path + SaveEnvStatsCallback,
        utils/callbacks.py:26-37)."""
        torch.save({"actor": self.actor.state_dict(), "critics": self.critics.state_dict(),
                    "actor_target": self.actor_target.state_dict(), "critics_target": self.critics_target.state_dict(),
                    "vec_normalize": self.vec_normalize.state_dict() if self.vec_normalize is not None else None,
                    "num_timesteps": self.num_timesteps, "n_updates": self.n_updates}, path)

    def load(self, path: str):
        sd = torch.load(path, map_location=self.device)
        self.actor.load_state_dict(sd["actor"]); self.critics.load_state_dict(sd["critics"])
        self.actor_target.load_state_dict(sd["actor_target"]); self.critics_target.load_state_dict(sd["critics_target"])
        if sd["vec_normalize"] is not None and self.vec_normalize is not None:
            self.vec_normalize.load_state_dict(sd["vec_normalize"])
        self.num_timesteps, self.n_updates = sd["num_timesteps"], sd["n_updates"]
        return self

    def close(self):
        """Drop the captured update graph (it holds NCCL work when gradients are averaged across ranks: the
        process group must not be destroyed under it)."""
        if self._graph is not None or self._pi_graphs:
            torch.cuda.synchronize(self.device)
            self._graph = None
            self._pi_graphs.clear()
            import gc
            gc.collect()
            torch.cuda.synchronize(self.device)

    def time_split(self):
        t = self.timers
        tot = max(t.env_step + t.heuristic + t.policy_forward + t.update, 1e-12)
        return {"env_step_s": t.env_step, "heuristic_s": t.heuristic, "policy_forward_s": t.policy_forward,
                "update_s": t.update, "env_step_frac": t.env_step / tot, "update_frac": t.update / tot}
