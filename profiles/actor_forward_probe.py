"""How long does the TD3 actor forward (4,096 x 28 -> 768 -> 4, tanh, unscale) take eagerly and as a graph replay?"""
import sys, time
import torch
sys.path.insert(0, ".")
from multi_echelon_drl_b200 import register_envs as R
from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
from multi_echelon_drl_b200.vec_normalize import VecNormalize

env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=4096, dtype="float32", seed=0), clip_obs=100.0, clip_reward=1000.0)
obs = env.reset()
for graph in (False, True):
    m = HgeTD3(env, None, TD3Config(cuda_graph=graph))
    for _ in range(10):
        m._actor_forward(obs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(1000):
        m._actor_forward(obs)
    b.record(); torch.cuda.synchronize()
    print(f"graph={graph}: GPU {a.elapsed_time(b):.1f} us/call, wall {(time.perf_counter() - t0) * 1e3:.1f} us/call, graphs captured: {len(m._pi_graphs)}")
    # one kernel-level look
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            m._actor_forward(obs)
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
        print(f"   {e.device_time_total / 20:8.1f} us/call  x{e.count // 20}  {e.key[:90]}")
