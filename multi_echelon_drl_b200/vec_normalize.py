"""Device-side VecNormalize for SupplyNetVecEnv.

Counterpart of `VecNormalize(make_vec_env(...), clip_obs=100, clip_reward=1000)` in the reference's
trainers (train_td3.py:100-101, train_a2c.py:86-87, train_dqn.py:87-88) with the running statistics,
the discounted-return accumulators and the normalisation itself on the GPU (`sn_vecnorm_update`,
`sn_vecnorm_apply`), so observations and rewards never leave the device.  `get_original_obs()` /
`get_original_reward()` return the un-normalised tensors that the replay buffer and the HGE heuristic
consume (hge.py:125); `save`/`load` persist the statistics like SaveEnvStatsCallback does
(utils/callbacks.py:26-37).

stable-baselines3 is third-party and absent here: semantics follow its documented behaviour
(SURVEY Appendix D); parity with SB3 itself is unpinned, the tests check against a numpy restatement.
"""
from __future__ import annotations

import ctypes as C
import pickle

import torch

from . import _lib
from .vec_env import LazyInfos


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecNormalize:
    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                 gamma=0.99, epsilon=1e-8):
        if getattr(venv, "output", "torch") != "torch":
            raise ValueError("VecNormalize runs on the device: build the env with output='torch'")
        self.venv = venv
        self.lib = _lib.load()
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = float(clip_obs), float(clip_reward), float(gamma), float(epsilon)
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.device = venv.device
        n, od, dev = self.num_envs, venv.spec.obs_dim, self.device
        self.od = od
        f64 = dict(dtype=torch.float64, device=dev)
        # RunningMeanStd(epsilon=1e-4): mean 0, var 1, count 1e-4.  ONE buffer each with a fixed address: the eager
        # kernels update it in place, and so does the step kernel when the update is fused into a replayed graph
        self._obs_rms = self._fresh(od, f64)
        self._ret_rms = self._fresh(1, f64)
        self.returns = torch.zeros(n, **f64)
        self._scratch = torch.zeros(2 * od + 2, **f64)
        # owned by the fused paths: sums that are zero between launches (two-kernel mode); one slot per block (one-kernel mode)
        self._fused_scratch = torch.zeros(2 * od + 2, **f64)
        self._fused_slots = torch.zeros(_lib.SN_VN_ONEPASS_SCRATCH_DOUBLES, **f64)
        self.old_obs = torch.zeros((n, od), dtype=torch.float32, device=dev)
        self.old_reward = torch.zeros((n,), dtype=torch.float32, device=dev)
        self._nobs = torch.zeros((n, od), dtype=torch.float32, device=dev)
        self._nrew = torch.zeros((n,), dtype=torch.float32, device=dev)
        self._nterm = torch.zeros((n, od), dtype=torch.float32, device=dev)
        if getattr(venv, "use_graph", False) and hasattr(venv, "attach_vecnorm"):
            venv.attach_vecnorm(self)

    @staticmethod
    def _fresh(d, kw):
        t = torch.zeros(2 * d + 1, **kw)
        t[d:2 * d] = 1.0
        t[2 * d] = 1e-4
        return t

    # ------------------------------------------------------------------ statistics access
    @property
    def obs_rms(self):
        r = self._obs_rms
        return {"mean": r[: self.od], "var": r[self.od: 2 * self.od], "count": r[2 * self.od]}

    @property
    def ret_rms(self):
        r = self._ret_rms
        return {"mean": r[0], "var": r[1], "count": r[2]}

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _update(self, obs, reward):
        rc = self.lib.sn_vecnorm_update(self.num_envs, self.od, _p(obs), _p(reward), self.gamma, _p(self.returns),
                                        _p(self._obs_rms), _p(self._obs_rms), _p(self._ret_rms), _p(self._ret_rms),
                                        _p(self._scratch), self._stream())
        _lib.check(rc)

    def frozen_stats(self):
        """Copy of the current statistics in buffers of their own: what a captured CUDA graph of the TD3 update
        normalises replay samples with while the live statistics move on; refresh by calling this again."""
        if not hasattr(self, "_frozen"):
            self._frozen = (torch.empty_like(self._obs_rms), torch.empty_like(self._ret_rms))
        self._frozen[0].copy_(self._obs_rms)
        self._frozen[1].copy_(self._ret_rms)
        return self._frozen

    def _apply(self, obs, obs_out, reward, reward_out, reset_returns, n=None, stats=None):
        obs_rms, ret_rms = stats if stats is not None else (self._obs_rms, self._ret_rms)
        rc = self.lib.sn_vecnorm_apply(self.num_envs if n is None else n, self.od, _p(obs), _p(obs_out), _p(reward), _p(reward_out),
                                       _p(obs_rms), _p(ret_rms), self.clip_obs, self.clip_reward,
                                       self.epsilon, int(reset_returns), _p(self.returns), self._stream())
        _lib.check(rc)

    # ------------------------------------------------------------------ joining the wrapped env's step graph
    def graph_key(self):
        """(fuse the statistics update into the step kernel, normalise inside the graph, one-kernel mode).  Fusion
        needs what the kernel supports: observations normalised (their moments come from the step's own tile),
        single-item beer-game chain; anything else keeps the eager kernels after the replay.  One-kernel mode: at most
        one 128-env tile per SM and 4-byte quantities - then the step launch normalises too (grid barrier between the
        moments and the normalisation) and the graph is that single kernel."""
        tag = (self.training, self.norm_obs, self.norm_reward, getattr(self, "no_onepass", False))
        cached = getattr(self, "_key_cache", None)
        if cached is not None and cached[0] == tag:          # called once per step: keep it to a tuple compare
            return cached[1]
        venv = self.venv
        spec, sim = venv.spec, venv.sim
        can_fuse = self.norm_obs and not spec.is_tree and spec.n_items <= 1 and sim.state_words == _lib.SN_STATE_WORDS
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        onepass = can_fuse and sim.dtype != "float64" and self.num_envs <= sms * 128 and not tag[3]
        key = (bool(self.training and can_fuse), bool(can_fuse or not self.training), bool(onepass))
        self._key_cache = (tag, key)
        return key

    def fused_args(self, key=None):
        update, _, onepass = key if key is not None else self.graph_key()
        d = dict(obs_rms=self._obs_rms, ret_rms=self._ret_rms, gamma=self.gamma,
                 scratch=(self._fused_slots if onepass else self._fused_scratch) if update else None,
                 returns=self.returns if (update and self.norm_reward) else None)
        if onepass:
            d.update(nobs=self._nobs, nrew=self._nrew if self.norm_reward else None, clip_obs=self.clip_obs,
                     clip_reward=self.clip_reward, epsilon=self.epsilon)
        return d

    def graph_tail(self):
        """Captured right after the step launch: normalise the step's outputs with the (just updated) statistics."""
        venv = self.venv
        self._apply(venv._obs if self.norm_obs else None, self._nobs, venv.sim.reward if self.norm_reward else None,
                    self._nrew, False)

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        obs = self.venv.reset()
        self.old_obs = obs
        self.returns.zero_()
        if not self.norm_obs:
            return obs
        if self.training:
            self._update(obs, None)
        self._apply(obs, self._nobs, None, None, False)
        return self._nobs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rew, dones, infos = self.venv.step_wait()
        # the wrapped env hands out the same two device buffers at every step: keep references, not copies
        self.old_obs, self.old_reward = obs, rew
        if getattr(self.venv, "last_step_graphed", False) and self.venv._graph_key[1]:
            # statistics update (if training) and normalisation ran inside the replayed graph
            return (self._nobs if self.norm_obs else obs), (self._nrew if self.norm_reward else rew), dones, infos
        done = infos.terminal_observation is not None
        if self.training:
            if self.norm_obs:
                # obs statistics and (if rewards are normalised) the discounted returns in one pass
                self._update(obs, rew if self.norm_reward else None)
            elif self.norm_reward:
                self._update_returns_only(rew)
        out_obs, out_rew = obs, rew
        self._apply(obs if self.norm_obs else None, self._nobs, rew if self.norm_reward else None, self._nrew,
                    done and self.norm_reward)
        if self.norm_obs:
            out_obs = self._nobs
        if self.norm_reward:
            out_rew = self._nrew
        elif done:
            self.returns.zero_()
        if done and self.norm_obs:
            # SB3 normalises infos[i]["terminal_observation"] with the statistics of this step
            self._apply(infos.terminal_observation, self._nterm, None, None, False)
            infos = LazyInfos(infos.n, self._nterm, infos.episode_return, infos.episode_length, infos.elapsed)
        return out_obs, out_rew, dones, infos

    def _update_returns_only(self, rew):
        """norm_obs=False, norm_reward=True: only the return statistics move (the obs statistics pass through a
        throw-away copy)."""
        keep = self._obs_rms.clone()
        self._update(self.old_obs, rew)
        self._obs_rms.copy_(keep)

    def step(self, actions):
        venv = self.venv
        if getattr(venv, "use_graph", False) and venv._graph is not None and actions is venv._act_buf:
            # fast path of the replayed graph (the general path below does the same with more Python in between)
            sim = venv.sim
            if not sim.terminal and sim.period + 1 < sim.T and venv._graph_key[1] and venv._graph_key == self.graph_key():
                venv._graph.replay()
                sim.period += 1
                venv.last_step_graphed = True
                self.old_obs, self.old_reward = venv._obs, sim.reward
                return ((self._nobs if self.norm_obs else venv._obs), (self._nrew if self.norm_reward else sim.reward),
                        venv._all_false, venv._no_infos)
        self.step_async(actions)
        return self.step_wait()

    def get_original_obs(self):
        """Un-normalised observation of the last step: the wrapped env's own output buffer, valid until the next
        step (clone it to keep it; SB3 returns a copy)."""
        return self.old_obs

    def get_original_reward(self):
        """Un-normalised reward of the last step (same lifetime as get_original_obs)."""
        return self.old_reward

    def normalize_obs(self, obs: torch.Tensor, stats=None) -> torch.Tensor:
        """Normalise any batch [m, obs_dim] with the current statistics (no update), or with `stats` =
        `frozen_stats()`."""
        if not self.norm_obs:
            return obs
        out = torch.empty_like(obs, dtype=torch.float32)
        self._apply(obs.to(torch.float32).contiguous(), out, None, None, False, n=obs.shape[0], stats=stats)
        return out

    def normalize_reward(self, reward: torch.Tensor, stats=None) -> torch.Tensor:
        if not self.norm_reward:
            return reward
        out = torch.empty_like(reward, dtype=torch.float32)
        self._apply(None, None, reward.to(torch.float32).contiguous(), out, False, n=reward.shape[0], stats=stats)
        return out

    def close(self):
        self.venv.close()

    def seed(self, seed=None):
        return self.venv.seed(seed)

    def __getattr__(self, name):          # spec, sim, period, ... of the wrapped env
        if name == "venv":
            raise AttributeError(name)
        return getattr(self.venv, name)

    # ------------------------------------------------------------------ persistence
    def state_dict(self):
        return {"obs_rms": self._obs_rms.cpu(), "ret_rms": self._ret_rms.cpu(), "clip_obs": self.clip_obs,
                "clip_reward": self.clip_reward, "gamma": self.gamma, "epsilon": self.epsilon,
                "norm_obs": self.norm_obs, "norm_reward": self.norm_reward}

    def load_state_dict(self, sd):
        self._obs_rms.copy_(sd["obs_rms"])
        self._ret_rms.copy_(sd["ret_rms"])
        for k in ("clip_obs", "clip_reward", "gamma", "epsilon", "norm_obs", "norm_reward"):
            setattr(self, k, sd[k])

    def save(self, save_path: str):
        with open(save_path, "wb") as fh:
            pickle.dump(self.state_dict(), fh)

    @staticmethod
    def load(load_path: str, venv):
        with open(load_path, "rb") as fh:
            sd = pickle.load(fh)
        vn = VecNormalize(venv)
        vn.load_state_dict(sd)
        return vn
