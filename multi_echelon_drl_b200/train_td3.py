"""Command line of the reference's train_td3.py (train_td3.py:18-147) on the GPU VecEnv.

    python -m multi_echelon_drl_b200.train_td3 -g 1 -p setup.yaml --name run --ordering-rule a \
        --role MultiFacility --scenario complex [--n-envs 4096] [--timesteps N]

Flags -g/-p/--name/--ordering-rule/--role/--scenario keep the reference's meaning.  Under torchrun every
rank owns n-envs envs (sharded by rank) and gradients are averaged over NCCL.
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import yaml

from .register_envs import make
from .td3 import HgeTD3, TD3Config
from .utils.heuristics import BaseStockPolicy
from .utils.wrappers import wrap_action_d_plus_a
from .vec_normalize import VecNormalize

ROLES = ["Retailer", "Wholesaler", "Distributor", "Manufacturer", "MultiFacility"]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-g", "--global-info", type=bool, default=False)     # any non-empty string is True (Q5)
    ap.add_argument("-p", "--hyperparameters", "--parameters", dest="parameters", type=str, default=None,
                    help="Path to the experiment setup file (.yaml); default: the packaged example for the role")
    ap.add_argument("--name", type=str, default="td3")
    ap.add_argument("--ordering-rule", type=str, default="a", choices=["a", "d+a"])
    ap.add_argument("--role", type=str, default="MultiFacility", choices=ROLES)
    ap.add_argument("--scenario", type=str, default="complex", choices=["basic", "complex"])
    ap.add_argument("--n-envs", type=int, default=4096)
    ap.add_argument("--timesteps", type=int, default=None)
    ap.add_argument("--log-interval", type=int, default=320)
    ap.add_argument("--no-cuda-graph", action="store_true", help="run the TD3 update eagerly instead of as a CUDA graph")
    args = ap.parse_args(argv)

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))

    # default setup = the example YAML that fits the role (the reference requires -p)
    here = os.path.dirname(os.path.abspath(__file__))
    path = args.parameters or os.path.join(here, "setup_example_centralized.yaml" if args.role == "MultiFacility"
                                           else "setup_example_decentralized.yaml")
    with open(path) as fh:
        setup = yaml.safe_load(fh)
    p = setup.get("hyperparameters", {}).get("td3", {})
    if args.scenario == "basic":                      # train_td3.py:52-61
        targets, demand_type, (lb, ub), offset = [48, 43, 41, 30], "Normal", (0, 20), -10
    else:
        targets, demand_type, (lb, ub), offset = [19, 20, 20, 14], "Uniform", (0, 16), -8
    if p.get("hge_rate_at_start", 0) > 0 and args.role != "MultiFacility":
        raise NotImplementedError("For now the TD3 with HGE only supports centralized setting")    # train_td3.py:65-66
    # the centralized ids exist as ...MultiFacilityFullInfo-v0 only (register_envs.py:74-110; the reference's docs always
    # pass -g 1 there): -g is implied for MultiFacility, so the flag-less default is a registered id
    full = bool(args.global_info) or args.role == "MultiFacility"
    env_name = f"BeerGame{demand_type}{args.role}{'FullInfo' * full}-v0"
    bsp = BaseStockPolicy(targets, {"on_hand": 0, "unreceived_pipeline": [3, 4, 5, 6], "unfilled_demand": 1,
                                    "latest_demand": 2}, 7, lb, ub, args.ordering_rule)
    env = make(env_name, n_envs=args.n_envs, dtype="float32", seed=1000 * rank, use_graph=True)
    if args.ordering_rule == "d+a":
        env = wrap_action_d_plus_a(env, offset=-(ub - lb) / 2, lb=lb, ub=ub)  # train_td3.py:86-91
    env = VecNormalize(env, clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(batch_size=p.get("batch_size", 128),
                    net_arch=(p.get("network_width", p.get("width", 256)) * p.get("num_layers", 3),),
                    learning_rate=p.get("learning_rate", 1e-3), tau=p.get("tau", 0.01), train_freq=p.get("train_freq", 32),
                    gradient_steps=p.get("gradient_steps", p.get("train_freq", 32)),
                    action_noise_std=p.get("action_noise_std", 0.4), gamma=p.get("gamma", 0.95),
                    hge_rate_at_start=p.get("hge_rate_at_start", 0.5), seed=rank,
                    cuda_graph=not args.no_cuda_graph)
    heuristic = bsp if cfg.hge_rate_at_start > 0 else None
    model = HgeTD3(env, heuristic, cfg, ddp=world > 1)
    total = args.timesteps or setup.get("max_time_steps", 10_000_000)
    import time
    t0 = time.time()
    model.learn(total, log_interval=args.log_interval if rank == 0 else None)
    torch.cuda.synchronize()
    wall = time.time() - t0
    if rank == 0:
        out = {"name": args.name, "env": env_name, "n_envs_per_gpu": args.n_envs, "n_gpus": world,
               "timesteps": model.num_timesteps, "updates": model.n_updates, "cuda_graph": cfg.cuda_graph,
               "wall_s": wall, "env_steps_per_s": model.num_timesteps * world / wall,
               "time_split": model.time_split(),
               "last_mean_episode_return": model.episode_returns[-1] if model.episode_returns else None}
        print(json.dumps(out))
    model.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
