"""Period step at 4 Mi envs on chains with other lead times (57 state words) and on a distribution tree (64 words): plain
vs TMA-tiled kernel, CUDA graph of 40 periods, observation out."""
import ctypes as C, json, os, sys
from dataclasses import replace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200 import _lib, uniform_spec
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
from multi_echelon_drl_b200.spec import tree_spec
PEAK, T, n = 6462.1, 100, 4 * 1024 * 1024
base = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
tree = tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20), dict(name="dist_b", demand=(0, 5), target=14),
                  dict(name="maker", target=25)],
                 [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)],
                 ["shop", "dist_a", "dist_b", "maker"], True)
specs = {"leadtimes": (replace(base, info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1)), 57), "tree": (tree, 64)}
os.environ["SN_NO_SPECIALIZED"] = "1"
out = {}
for name, (sp, words) in specs.items():
    sim = BatchedSupplyNetwork(sp, n, "cuda", "int32")
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device="cuda", dtype=torch.int32))
    sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device="cuda", dtype=torch.int32))
    acts = torch.randint(0, 17, (n, 4), generator=g, device="cuda", dtype=torch.int32)
    for which in ("plain", "tiled"):
        os.environ["SN_STEP_KERNEL"] = which
        sim.reset()
        def one(t):
            _lib.check(sim.lib.sn_step(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, t, C.c_void_p(sim.state.data_ptr()),
                                       C.c_void_p(acts.data_ptr()), C.c_void_p(sim.demand[t + 1].data_ptr()),
                                       C.c_void_p(sim.obs.data_ptr()), C.c_void_p(sim.reward.data_ptr()), None, None, None, None,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for t in range(3): one(t)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for t in range(3, 43): one(t)
        gr.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3): gr.replay()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 120
        bps = 2 * words * 4 + 16 + 4 * sp.n_sources + 4 + 112
        out[f"{name}_{which}"] = {"us": round(ms * 1e3, 1), "frac": round(n * bps / (ms * 1e-3) / 1e9 / PEAK, 4), "bytes_per_env_step": bps}
    del sim
print(json.dumps(out))
