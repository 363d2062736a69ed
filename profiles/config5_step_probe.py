"""Config-5 env step (4,096 envs, float32, VecNormalize fused) broken down: graphs of 40 identical device-period launches."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200.register_envs import make
from multi_echelon_drl_b200.vec_normalize import VecNormalize
out = {}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TAGS = ("step_plain", "step_tiled", "step_tiled_vn", "step_tiled_vn+apply", "apply_only", "onepass", "onepass_eval")
if os.environ.get("SN_PROBE_TAGS"):                      # e.g. SN_PROBE_TAGS=onepass under ncu
    TAGS = tuple(os.environ["SN_PROBE_TAGS"].split(","))
for tag in TAGS:
    env = VecNormalize(make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=0, use_graph=True),
                       clip_obs=100.0, clip_reward=1000.0)
    env.reset()
    venv, sim = env.venv, env.venv.sim
    venv.action_buffer.fill_(4.0)
    sim._graph_period_valid = True
    os.environ["SN_STEP_KERNEL"] = "plain" if tag == "step_plain" else "tiled"
    env.training = tag != "onepass_eval"
    key = (env.training, True, tag.startswith("onepass"))
    def body():
        if tag != "apply_only":
            sim.launch_step_device_period(venv.action_buffer, venv._obs, sim.reward,
                                          vecnorm=env.fused_args(key) if ("vn" in tag or "onepass" in tag) else None)
        if "apply" in tag:
            env.graph_tail()
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(40): body()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.reset(venv._obs); torch.cuda.synchronize()
    s.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
    out[tag] = round(s.elapsed_time(e) / 80 * 1e3, 2)
    assert sim.device_flags() == 0, tag
print(json.dumps(out))
