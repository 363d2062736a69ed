"""Period-step kernel A/B on a B200: plain (one thread per env, scalar loads) vs TMA-tiled persistent kernel, at
the BASELINE sizes and at an HBM-honest 4 Mi envs; multi-item step (config 4) eager vs graph; the config-5 env step
(4,096 envs + VecNormalize) eager vs one replayed graph.  usage: python profiles/step_tiled_probe.py [tag]
SUPPLYNET_LIB selects a library variant (e.g. built with -DSN_TILED_STAGES=3)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multi_echelon_drl_b200 import _lib, uniform_spec  # noqa: E402
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork  # noqa: E402

PEAK = 6462.1
T = 100
out = {"lib": os.environ.get("SUPPLYNET_LIB", "default"), "tag": sys.argv[1] if len(sys.argv) > 1 else ""}
dev = torch.device("cuda")
spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)


def time_graph(fn, reps=5):
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        s.record()
        for _ in range(reps):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / reps)
    return best


for n in (4096, 65536, 131072, 262144, 524288, 4 * 1024 * 1024):
    sim = BatchedSupplyNetwork(spec, n, dev, "int32")
    g = torch.Generator(device=dev); g.manual_seed(1)
    sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device=dev, dtype=torch.int32))
    sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device=dev, dtype=torch.int32))
    acts = torch.randint(0, 17, (n, 4), generator=g, device=dev, dtype=torch.int32)
    for which in ("plain", "tiled"):
        os.environ["SN_STEP_KERNEL"] = which
        for with_obs in (False, True):
            sim.reset()

            def one(t):
                _lib.check(sim.lib.sn_step(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, t, C.c_void_p(sim.state.data_ptr()),
                                           C.c_void_p(acts.data_ptr()), C.c_void_p(sim.demand[t + 1].data_ptr()),
                                           C.c_void_p(sim.obs.data_ptr()) if with_obs else None,
                                           C.c_void_p(sim.reward.data_ptr()), None, None, None, None,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            for t in range(3):
                one(t)
            torch.cuda.synchronize()
            steps = 40
            ms = time_graph(lambda: [one(t) for t in range(3, 3 + steps)]) / steps
            bps = 344 + (112 if with_obs else 0)
            out[f"step_{n}_{which}_{'obs' if with_obs else 'state'}"] = {
                "us": round(ms * 1e3, 3), "gbs": round(n * bps / (ms * 1e-3) / 1e9, 1), "frac": round(n * bps / (ms * 1e-3) / 1e9 / PEAK, 4)}
    os.environ.pop("SN_STEP_KERNEL", None)
    del sim

# config 4: 262,144 envs x 2 items
from dataclasses import replace  # noqa: E402
from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork  # noqa: E402
n = 262144
specs = [spec, replace(spec, holding_cost=(0.75,) * 4)]
for use_graph in (False, True):
    mi = MultiItemSupplyNetwork(specs, n, dev, "int32", use_graph=use_graph)
    g = torch.Generator(device=dev); g.manual_seed(2)
    mi.sim.demand.random_(0, 9, generator=g)
    mi.sim.init.random_(0, 9, generator=g)
    acts = torch.randint(0, 17, (n, 8), generator=g, device=dev, dtype=torch.int32)
    mi.action_buffer.copy_(acts)
    a = mi.action_buffer if use_graph else acts
    mi.reset()
    for _ in range(5):
        mi.step(a)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    reps = 60
    for _ in range(reps):
        mi.step(a)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    bpc = 344 + 112 + 2 + 16   # per chain: state r/w 320, action 16, demand 4, reward 4, obs 112, env reward 2 (one f32 per 2 chains), episode return r/w 16
    out[f"multi_item_262144x2_{'graph' if use_graph else 'eager'}"] = {
        "us": round(us, 2), "gbs": round(2 * n * bpc / (us * 1e-6) / 1e9, 1), "frac": round(2 * n * bpc / (us * 1e-6) / 1e9 / PEAK, 4)}
    del mi

# config 5 env step: 4,096 envs, VecNormalize(clip 100 / 1000), float32
from multi_echelon_drl_b200.register_envs import make  # noqa: E402
from multi_echelon_drl_b200.vec_normalize import VecNormalize  # noqa: E402
for use_graph in (False, True):
    env = VecNormalize(make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=4096, dtype="float32", seed=0, use_graph=use_graph),
                       clip_obs=100.0, clip_reward=1000.0)
    env.reset()
    a = env.action_buffer
    a.fill_(4.0)
    for _ in range(10):
        env.step(a)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    reps = 80                      # stays inside one episode: no terminal step in the timed region
    for _ in range(reps):
        env.step(a)
    e.record()
    torch.cuda.synchronize()
    out[f"config5_env_step_{'graph' if use_graph else 'eager'}"] = {
        "gpu_us": round(s.elapsed_time(e) / reps * 1e3, 2), "wall_us": round((time.perf_counter() - t0) / reps * 1e6, 2)}
print(json.dumps(out))
