"""Does centralized TD3+HGE learn on the GPU VecEnv (config 5)?  Trains a few variants for a bounded number of
vector steps and prints the training-episode returns and a deterministic evaluation, next to the base-stock
heuristic and the uniform-random policy on the same env.  usage: python profiles/td3_learning_probe.py [variant ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_echelon_drl_b200.register_envs import make
from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
from multi_echelon_drl_b200.vec_normalize import VecNormalize

ENV = "BeerGameUniformMultiFacilityFullInfo-v0"
VARIANTS = {
    # the reference's hyper-parameters (setup_exmaple_centralized.yaml) with the HGE decay stretched to the run
    "ref": dict(n=4096, episodes=40, batch_size=128, lr=1e-3, tau=0.01, noise=0.4, gsteps=32, decay=0.5),
    "b1024": dict(n=4096, episodes=40, batch_size=1024, lr=1e-3, tau=0.01, noise=0.4, gsteps=32, decay=0.5),
    "b1024_g128": dict(n=4096, episodes=40, batch_size=1024, lr=1e-3, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "b1024_lr3e4": dict(n=4096, episodes=40, batch_size=1024, lr=3e-4, tau=0.01, noise=0.3, gsteps=64, decay=0.5),
    # robustness sweep over seeds (SN_PROBE_SEEDS=0,1,2,3): which recipe learns on EVERY seed
    "g128_lr3e4": dict(n=4096, episodes=40, batch_size=1024, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "g128_lr3e4_b2048": dict(n=4096, episodes=40, batch_size=2048, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "g128_tau005": dict(n=4096, episodes=40, batch_size=1024, lr=1e-3, tau=0.005, noise=0.3, gsteps=128, decay=0.5),
    "g128_noise02": dict(n=4096, episodes=40, batch_size=1024, lr=1e-3, tau=0.01, noise=0.2, gsteps=128, decay=0.5),
    "g128_lr3e4_e60": dict(n=4096, episodes=60, batch_size=1024, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "lr3e4_b2048_e60": dict(n=4096, episodes=60, batch_size=2048, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "lr3e4_b2048_e80": dict(n=4096, episodes=80, batch_size=2048, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "lr3e4_b2048_e80_d35": dict(n=4096, episodes=80, batch_size=2048, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.35),
    "lr3e4_b2048_g256": dict(n=4096, episodes=40, batch_size=2048, lr=3e-4, tau=0.01, noise=0.3, gsteps=256, decay=0.5),
    "lr5e4_b2048": dict(n=4096, episodes=40, batch_size=2048, lr=5e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "lr3e4_b4096": dict(n=4096, episodes=40, batch_size=4096, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5),
    "g128_buf4m": dict(n=4096, episodes=40, batch_size=1024, lr=1e-3, tau=0.01, noise=0.3, gsteps=128, decay=0.5, buffer=4_000_000),
    "g128_lr3e4_buf4m": dict(n=4096, episodes=40, batch_size=1024, lr=3e-4, tau=0.01, noise=0.3, gsteps=128, decay=0.5, buffer=4_000_000),
}


def baseline_returns(n=4096):
    env = make(ENV, n_envs=n, dtype="float32", seed=7)
    pol = BaseStockPolicy([19, 20, 20, 14])
    obs = env.reset()
    out = {}
    for name in ("heuristic", "random"):
        rets = None
        while rets is None:
            a = pol.get_order_quantity(obs) if name == "heuristic" else torch.rand((n, 4), device="cuda") * 16
            obs, _, d, info = env.step(a)
            if info.terminal_observation is not None:
                rets = info.episode_return.mean().item()
        out[name] = rets
    return out


def run(name, v, seed=None):
    n = v["n"]
    seed = int(os.environ.get("SN_PROBE_SEED", "0")) if seed is None else seed
    env = VecNormalize(make(ENV, n_envs=n, dtype="float32", seed=0, use_graph=True), clip_obs=100.0, clip_reward=1000.0)
    total = n * 100 * v["episodes"]
    cfg = TD3Config(batch_size=v["batch_size"], net_arch=(768,), learning_rate=v["lr"], tau=v["tau"], train_freq=32,
                    gradient_steps=v["gsteps"], action_noise_std=v["noise"], gamma=0.95, hge_rate_at_start=0.5,
                    hge_decay_periods=max(1, int(total * v["decay"] / 100)), learning_starts=n * 100, seed=seed,
                    buffer_size=v.get("buffer", 1_000_000))
    model = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), cfg)
    t0 = time.time()
    model.learn(total)
    torch.cuda.synchronize()
    wall = time.time() - t0
    eval_env = VecNormalize(make(ENV, n_envs=n, dtype="float32", seed=123), clip_obs=100.0, clip_reward=1000.0)
    mean, std, cnt = model.evaluate(eval_env, n_episodes=1)
    res = {"variant": name, "seed": seed, **v, "wall_s": round(wall, 1), "updates": model.n_updates,
           "train_returns": [round(r) for r in model.episode_returns], "eval_mean": round(mean, 1), "eval_std": round(std, 1)}
    model.close()
    return res


if __name__ == "__main__":
    print(json.dumps({"baselines": baseline_returns()}), flush=True)
    seeds = os.environ.get("SN_PROBE_SEEDS")
    for name in (sys.argv[1:] or list(VARIANTS)):
        if seeds:
            evals = []
            for sd in (int(x) for x in seeds.split(",")):
                r = run(name, VARIANTS[name], sd)
                evals.append(r["eval_mean"])
                print(json.dumps({"variant": name, "seed": sd, "last_train": r["train_returns"][-3:], "eval_mean": r["eval_mean"],
                                  "wall_s": r["wall_s"]}), flush=True)
            print(json.dumps({"variant": name, "evals": evals, "worst": min(evals)}), flush=True)
        else:
            print(json.dumps(run(name, VARIANTS[name])), flush=True)
