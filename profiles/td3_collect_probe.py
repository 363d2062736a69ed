"""Where a config-5 vector step spends its time: collection alone (train_freq 32, no updates), updates alone
(train(32) on the filled buffer), both; wall clock and GPU time.  usage: python profiles/td3_collect_probe.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multi_echelon_drl_b200.register_envs import make  # noqa: E402
from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config  # noqa: E402
from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy  # noqa: E402
from multi_echelon_drl_b200.vec_normalize import VecNormalize  # noqa: E402

n = 4096
out = {}
for graph_env in (True,):
    env = VecNormalize(make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=0, use_graph=graph_env),
                       clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(batch_size=128, net_arch=(768,), train_freq=32, gradient_steps=32, hge_rate_at_start=0.5, seed=0)
    m = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), cfg)
    m.learn(n * 128)
    torch.cuda.synchronize()
    out["collect_graphs"] = m._collect is not None and m._collect.store is not None

    def timed(fn, reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        t1 = time.perf_counter()          # CPU done issuing
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return {"issue_us": round((t1 - t0) / reps * 1e6, 1), "wall_us": round((t2 - t0) / reps * 1e6, 1),
                "gpu_us": round(a.elapsed_time(b) / reps * 1e3, 1)}

    out["both_per_step"] = timed(lambda: m.learn(m.num_timesteps + n * 256), 256)
    out["both_per_step_again"] = timed(lambda: m.learn(m.num_timesteps + n * 256), 256)
    out["update_only_per_update"] = timed(lambda: m.train(256), 256)
    m.cfg.gradient_steps = 0
    out["collect_only_per_step"] = timed(lambda: m.learn(m.num_timesteps + n * 256), 256)
    m.hge_rate = 0.0
    cg = m._collect
    if cg is not None:
        out["pi_graph"] = timed(lambda: [cg.pi.replay() for _ in range(200)], 200)
        out["heur_graph"] = timed(lambda: [cg.heur.replay() for _ in range(200)], 200)
        keep_pos = m.buffer.pos_t.clone()
        out["store_graph"] = timed(lambda: [cg.store.replay() for _ in range(200)], 200)
        m.buffer.pos_t.copy_(keep_pos)
        out["env_graph"] = timed(lambda: [env.venv._graph.replay() for _ in range(50)], 50)
print(json.dumps(out))
