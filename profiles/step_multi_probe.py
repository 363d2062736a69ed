"""Where does the multi-item period step lose time?  524,288 chains, tiled kernel, variants of the optional outputs."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from multi_echelon_drl_b200 import _lib, uniform_spec
from multi_echelon_drl_b200.multi_item import combine_item_specs
from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
dev = torch.device("cuda"); T = 100; n = 524288
base = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
multi = combine_item_specs([base, replace(base, holding_cost=(0.75,) * 4)])
os.environ["SN_STEP_KERNEL"] = sys.argv[1] if len(sys.argv) > 1 else "tiled"
out = {}
for tag, spec, use_er, use_ep in (("single", base, 0, 0), ("single+ep", base, 0, 1), ("multi", multi, 0, 0),
                                  ("multi+envrew", multi, 1, 0), ("multi+envrew+ep", multi, 1, 1)):
    sim = BatchedSupplyNetwork(spec, n, dev, "int32")
    g = torch.Generator(device=dev); g.manual_seed(1)
    sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device=dev, dtype=torch.int32))
    sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device=dev, dtype=torch.int32))
    acts = torch.randint(0, 17, (n, 4), generator=g, device=dev, dtype=torch.int32)
    er = torch.zeros(n // 2, dtype=torch.float32, device=dev)
    sim.reset()
    def one(t):
        io = _lib.SnStepIo()
        io.state, io.actions, io.demand, io.period = sim.state.data_ptr(), acts.data_ptr(), sim.demand[t + 1].data_ptr(), t
        io.obs, io.reward = sim.obs.data_ptr(), sim.reward.data_ptr()
        if use_er: io.env_reward = er.data_ptr()
        if use_ep: io.ep_return = sim.ep_return.data_ptr()
        _lib.check(sim.lib.sn_step_ex(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, C.byref(io), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for t in range(3): one(t)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for t in range(3, 43): one(t)
    gr.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): gr.replay()
    e.record(); torch.cuda.synchronize()
    out[tag] = round(s.elapsed_time(e) / 200 * 1e3, 2)
    del sim
print(json.dumps(out))
