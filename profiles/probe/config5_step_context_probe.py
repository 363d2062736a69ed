"""Why does the one-kernel env step take 9.3 us in profiles/config5_step_probe.py and 10.6 us in bench.py's config-5
block?  Same launch, graph of 40, timed (a) on a fresh env with constant actions, (b) fresh env with random actions,
(c) after 128 TD3 vector steps (policy actions in the buffer, statistics of a real run), (d) the same with the action
buffer overwritten by the constant."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from multi_echelon_drl_b200.register_envs import make
from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
from multi_echelon_drl_b200.vec_normalize import VecNormalize

n = 4096
env = VecNormalize(make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=0, use_graph=True),
                   clip_obs=100.0, clip_reward=1000.0)
venv, sim = env.venv, env.venv.sim
out = {}


def timed(tag):
    env.reset()
    sim._graph_period_valid = True
    sim.ctl[0:1].fill_(sim.period)
    torch.cuda.synchronize()
    key = env.graph_key()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(40):
            sim.launch_step_device_period(venv.action_buffer, venv._obs, sim.reward, vecnorm=env.fused_args(key))
    best = 1e9
    for _ in range(3):
        env.reset()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 80 * 1e3)
    out[tag] = round(best, 2)
    assert sim.device_flags() == 0


env.reset()
venv.action_buffer.fill_(4.0)
timed("fresh_constant_actions")
venv.action_buffer.copy_(torch.rand_like(venv.action_buffer) * 16)
timed("fresh_random_actions")
model = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), TD3Config(batch_size=128, net_arch=(768,), train_freq=32, gradient_steps=32))
model.learn(n * 128)
torch.cuda.synchronize()
timed("after_td3_policy_actions")
venv.action_buffer.fill_(4.0)
timed("after_td3_constant_actions")
out["obs_abs_max"] = float(venv._obs.abs().max())
print(json.dumps(out))
