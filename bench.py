#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched beer game on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE config 2 — complex ("uniform") scenario, MultiFacility
centralized control with full information, 65,536 envs per GPU, fixed integer actions, 100-period
episodes.  One bench *step* = one full parity rollout of that batch: one sn_episode launch streaming
the actions [100, N, 4] and demands in and the complete trajectory (obs [100, N, 28] f32, reward
[100, N] f32) out = 6,553,600 env-steps and ~0.9 GB of HBM traffic per step (inputs + outputs are
several times the 126 MB L2, so no L2 flush is needed between steps).

  value   device-resident throughput: inputs already in HBM, outputs left in HBM
  e2e     the same rollout through the public Python API with HOST (pinned) buffers: init/demand/
          actions copied host->device and the trajectory copied device->host inside the timed region
  roofline  k_rollout: algorithmic bytes (137 B per env-step) / CUDA-event time of that
          kernel / measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the C port of the oracle (oracle/beergame_c.c) on the host cores, bounded sample

`--impl reference` times the CPU implementation of the same path (oracle port; the reference itself
is pure Python and is not present on the GPU box) with all host threads.
Multi-GPU: one process per GPU (torchrun), envs sharded by rank, no data-path collective; the only
collective is the NCCL all-reduce of the 16-double statistics vector per episode.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 65536
T = 100
# SURVEY 8(d) counts one more byte per env-step for `done`; all envs run in lockstep, `done` is a host bool and no
# kernel writes it, so it is not counted here
BYTES_PER_ENV_STEP_ROLLOUT = 136      # in: action 16 + demand 4; out: obs 112 + reward 4
BYTES_PER_ENV_STEP_SINGLE = 344       # single-step VecEnv mode (state 160 r + 160 w, action 16, demand 4, reward 4)
WORKLOAD = ("config2: complex (uniform) scenario, MultiFacility centralized full-info, {n} envs/GPU x {T} periods "
            "fixed-action parity rollout, full trajectory out")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:      # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def ncu_traffic(n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (config-2 size only)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if n != N_ENVS or not os.path.isfile(p):
        return None
    with open(p) as f:
        d = json.load(f)["k_rollout_config2"]
    return d["dram_bytes_read"] + d["dram_bytes_write"]


def make_inputs(n, seed):
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
    rs = np.random.RandomState(seed)
    init, demand = bulk_episode_inputs(spec, n, rs)
    actions = rs.randint(0, 17, (T, n, 4)).astype(np.int32)
    return spec, init.astype(np.int32), demand.astype(np.int32), actions


def cpu_port_throughput(init, demand, actions, threads, budget_s, with_traj=True):
    """C port of the oracle on `threads` host threads; repeats the 65,536-env rollout until
    ~budget_s seconds of work are done.  Returns (env-steps/s, sample description)."""
    from oracle import beergame_c as OC
    from oracle import beergame_np as O
    ospec = O.scenario_spec("uniform", (0, 1, 2, 3), True, T)
    n = init.shape[0]
    OC.rollout(ospec, init[:1024], demand[:1024], actions[:, :1024], n_threads=threads)      # warm-up
    lib = OC.load()
    import ctypes as C
    s = OC.make_spec(ospec)
    od = 28
    obs = np.zeros((T, n, od), np.float32) if with_traj else None
    rew = np.zeros((T, n), np.float64) if with_traj else None
    ret = np.zeros((n,), np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    reps, t0 = 0, time.perf_counter()
    while True:
        lib.bg_rollout(C.byref(s), n, T, p(init), p(demand), p(actions), None, threads, None, p(obs), p(rew), p(ret))
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= 200:
            break
    return n * T * reps / dt, f"{reps} x ({n} envs x {T} periods, trajectory {'written' if with_traj else 'discarded'}) in {dt:.1f}s"


def _py_ref_worker(args):
    """One worker of the reference's own Python stepping loop: 16 envs stepped round-robin,
    per-env np.random.seed before reset, fixed action [4,4,4,4]; timing inside the worker."""
    wid, n_envs, episodes, ref_root = args
    import warnings
    os.environ["SUPPLYNET_REFERENCE_ROOT"] = ref_root
    from oracle import ref_harness as H
    warnings.simplefilter("ignore")
    envs = [H.make_env("uniform", H.NODE_NAMES, True, True, 100) for _ in range(n_envs)]
    a = np.array([4, 4, 4, 4])
    t0 = time.perf_counter()
    steps = 0
    for ep in range(episodes):
        for i, e in enumerate(envs):
            np.random.seed(1000 * wid + 16 * ep + i)
            e.reset()
        for _ in range(100):
            for e in envs:
                e.step(a)
                steps += 1
    return steps, time.perf_counter() - t0


def python_reference_throughput(episodes=6):
    """The UNMODIFIED reference (make_beer_game_uniform_multi_facility + env.step, envs.py:289-414, 82-113) on the host
    cores: multiprocessing.Pool(P = os.cpu_count()) x 16 envs x `episodes` episodes, plus a 1-process run.  The
    reference travels to the GPU box as the git-ignored copy baseline/_ref (made by __graft_entry__.build() where
    /root/reference is mounted); without it this leg reports why it is absent."""
    import multiprocessing as mp
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_root, "supplynetwork.py")):
        ref_root = "/root/reference"
    if not os.path.isfile(os.path.join(ref_root, "supplynetwork.py")):
        return {"unavailable": "baseline/_ref (copy of the reference) is not present on this box"}
    P = os.cpu_count() or 1
    out = {"unit": "env-steps/s", "kind": "reference", "envs_per_worker": 16, "episodes": episodes,
           "source": os.path.relpath(ref_root, ROOT) if ref_root.startswith(ROOT) else ref_root,
           "workload": "make_beer_game_uniform_multi_facility(all four facilities, box_action_space=True, random_init=True, "
                       "global_observable=True), fixed action [4,4,4,4], per-env np.random.seed, resets included"}
    ctx = mp.get_context("spawn")                    # no CUDA context / pinned buffers inherited by the workers
    for procs in sorted({1, P}):
        with ctx.Pool(procs) as pool:
            res = pool.map(_py_ref_worker, [(w, 16, episodes, ref_root) for w in range(procs)])
        total, slowest = sum(r[0] for r in res), max(r[1] for r in res)
        key = "value" if procs == P else "single_process_value"
        out[key] = total / slowest
        if procs == P:
            out.update(cores=P, per_core_value=total / slowest / P,
                       sample=f"{procs} processes x 16 envs x {episodes} episodes x 100 periods = {total} env-steps, slowest worker {slowest:.1f}s")
        if P == 1:
            out["single_process_value"] = out["value"]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, init, demand, actions = make_inputs(N_ENVS, 1234)
    threads = os.cpu_count() or 1
    from oracle import beergame_c as OC
    OC.load()
    vals = []
    per_step_budget = min(20.0, 120.0 / max(1, args.steps + args.warmup))
    per_step_budget = float(os.environ.get("BENCH_CPU_BUDGET_S", per_step_budget))     # tests shrink the sample
    for i in range(args.warmup + args.steps):
        v, sample = cpu_port_throughput(init, demand, actions, threads, per_step_budget)
        if i >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    v1, _ = cpu_port_throughput(init, demand, actions, 1, min(3.0, per_step_budget))
    line = {
        "impl": "reference", "metric": "env-steps/sec (batched beer game)", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": N_ENVS * T / v * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=N_ENVS, T=T),
                   "implementation": "C port of the reference algorithm (oracle/beergame_c.c) on all host threads; the reference's own Python loop is timed beside it (cpu_baseline.python_reference)"},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample,
                         "single_thread_value": v1,
                         "note": "C port of the reference algorithm (oracle/beergame_c.c): the stronger CPU baseline, used as this line's value",
                         "python_reference": None if args.no_python_ref else python_reference_throughput()},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_native(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from multi_echelon_drl_b200.distributed import bind_to_gpu_cpus
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork

    # one process per GPU: stay on the CPUs / memory of the socket this GPU hangs off before pinning host buffers
    all_cpus = os.sched_getaffinity(0)
    cpus = None if args.no_numa_bind else bind_to_gpu_cpus(local)
    n = args.envs
    spec, init, demand, actions = make_inputs(n, 1234 + rank)
    sim = BatchedSupplyNetwork(spec, n, dev, "int32")
    sim.track_stats = True
    # host (pinned) buffers for the e2e arm
    h_init = torch.from_numpy(init).pin_memory()
    h_demand = torch.from_numpy(demand).pin_memory()
    h_actions = torch.from_numpy(actions).pin_memory()
    h_traj_obs = torch.empty((T, n, 28), dtype=torch.float32).pin_memory()
    h_traj_rew = torch.empty((T, n), dtype=torch.float32).pin_memory()
    # device-resident inputs / outputs
    sim.load_episode(h_init, h_demand)
    d_actions = h_actions.to(dev)
    traj_obs = torch.empty((T, n, 28), dtype=torch.float32, device=dev)
    traj_rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    stats_bufs = [torch.zeros(16, dtype=torch.float64, device=dev) for _ in range(2)]
    pending, step_no = [], [0]
    # the statistics all-reduce: one kernel over peer memory (sn_stats_allreduce) where the ranks can map each other's
    # mailboxes, NCCL otherwise; the JSON line says which ran
    peer_ar, peer_ar_why = None, None
    if world > 1 and not args.nccl_stats:
        try:
            from multi_echelon_drl_b200.distributed import PeerStatsAllReduce
            peer_ar = PeerStatsAllReduce(dev)
        except Exception as exc:          # noqa: BLE001  (no peer mapping on this box: NCCL carries the vector)
            peer_ar_why = f"{type(exc).__name__}: {exc}"[:200]
    per_rank_ms, local_ms = {}, {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k_events = []

    def device_step(record):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        sim.episode(T, actions=d_actions, traj_obs=traj_obs, traj_reward=traj_rew, write_state=False)
        if record:
            e1.record()
            k_events.append((e0, e1))
        step_no[0] += 1
        if world > 1 and step_no[0] % args.stats_interval == 0:
            # the path's only collective: the statistics vector (16 doubles) that the kernels keep
            # accumulating on every rank, all-reduced once per log interval (every --stats-interval
            # episodes, and once more at the end of the timed region): one kernel over peer memory on this stream,
            # or (fallback) NCCL, issued asynchronously on its stream with two alternating buffers.
            buf = stats_bufs[(step_no[0] // args.stats_interval) % 2]
            if peer_ar is not None:       # one small launch on this stream: stores to the peers, sum of what arrived
                peer_ar(sim.stats, buf)
                return
            while len(pending) >= 2:
                pending.pop(0).wait()
            buf.copy_(sim.stats)
            pending.append(dist.all_reduce(buf, async_op=True))

    def e2e_step():
        # host buffers in, host trajectory out; H2D / kernels / D2H pipelined over spans of periods
        sim.episode_host(h_init, h_demand, h_actions, h_traj_obs, h_traj_rew, chunks=args.e2e_chunks)

    def timed(fn, steps, warmup, **kw):
        for _ in range(warmup):
            fn(**kw) if kw else fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn(**kw) if kw else fn()
        if world > 1 and fn is device_step:           # end-of-interval reduction of the accumulated statistics
            if peer_ar is not None:
                peer_ar(sim.stats, stats_bufs[0])
            else:
                stats_bufs[0].copy_(sim.stats)
                pending.append(dist.all_reduce(stats_bufs[0], async_op=True))
        while pending:
            pending.pop(0).wait()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        local_ms[fn.__name__] = float(ms.item())
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank_ms[fn.__name__] = [round(float(x.item()) / steps, 5) for x in allms]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        device_step(False)
    if world > 1:
        # the statistics all-reduce is part of the steady state: warm it up too (the first collectives of a process
        # set up channels and load kernels, milliseconds that do not belong to a 50-step window)
        for b in stats_bufs * 2:
            if peer_ar is not None:
                peer_ar(sim.stats, b)
            else:
                b.copy_(sim.stats)
                dist.all_reduce(b, async_op=True).wait()
        step_no[0] = 0
    # the timed region holds the K launches and nothing else (an event pair around every launch keeps consecutive
    # launches from overlapping their launch latency with the previous kernel's tail: 4-6 us per 152 us step); one step
    # = one launch, so this rank's time / K is the kernel's average launch duration, gaps included
    ms_dev = timed(device_step, args.steps, 0, record=False)
    kernel_ms = local_ms["device_step"] / args.steps
    for _ in range(min(args.steps, 10)):                  # for the record: the same launches with an event pair each
        device_step(True)
    while pending:
        pending.pop(0).wait()
    torch.cuda.synchronize()
    kernel_ms_event_pairs = float(np.mean([a.elapsed_time(b) for a, b in k_events]))
    ms_e2e = timed(e2e_step, max(1, min(args.steps, 10)), min(args.warmup, 3))
    e2e_steps = max(1, min(args.steps, 10))
    clocks = sampler.finish() if sampler else None       # sampled across both timed regions

    # correctness guard on the numbers just produced: statistics vector vs the trajectory it wrote
    sim.stats.zero_()
    device_step(False)
    torch.cuda.synchronize()
    assert float(sim.stats[0].item()) == n * T, "stats vector: wrong env-step count"
    assert float(sim.stats[1].item()) == float(traj_rew.double().sum().item()), "stats vector disagrees with trajectory"

    # the peer-memory all-reduce against NCCL on the same vector (each rank contributes something different)
    stats_exchange = None
    if world > 1:
        if peer_ar is not None:
            probe = sim.stats.clone() * (1.0 + rank) + rank
            via_nccl = probe.clone()
            dist.all_reduce(via_nccl)
            via_peer = peer_ar(probe, torch.empty_like(probe))
            torch.cuda.synchronize()
            same = bool(torch.allclose(via_peer, via_nccl, rtol=1e-13, atol=0.0))
            assert same, "sn_stats_allreduce disagrees with NCCL"
            stats_exchange = {"how": "one kernel over peer memory (sn_stats_allreduce: stores into every rank's mailbox over NVLink, sum in rank order)",
                              "equals_nccl": same, "calls": peer_ar.seq}
        else:
            stats_exchange = {"how": "NCCL all-reduce", "why_not_peer_memory": peer_ar_why or "--nccl-stats"}

    # shard invariance across real devices: the envs of one global batch, sharded by index over the ranks,
    # give bit for bit what the whole batch gives on ONE GPU (rank 0 runs it and compares every env)
    shard_check = None
    if world > 1:
        m = 8192
        g_spec, g_init, g_demand, g_actions = make_inputs(m * world, 4321)
        lo = rank * m

        def per_env(sim_x, init_x, demand_x, actions_x, k):
            sim_x.load_episode(init_x, demand_x)
            to = torch.empty((T, k, 28), dtype=torch.float32, device=dev)
            sim_x.episode(T, actions=torch.as_tensor(actions_x).to(dev), traj_obs=to, write_state=False)
            # (position-weighted so that a permutation inside a trajectory would show)
            w = torch.arange(1, T * 28 + 1, dtype=torch.float64, device=dev).view(T, 1, 28)
            return torch.stack([(to.double() * w).sum((0, 2)), sim_x.ep_return.clone()])

        mine = per_env(BatchedSupplyNetwork(g_spec, m, dev, "int32"), g_init[lo:lo + m], g_demand[lo:lo + m],
                       g_actions[:, lo:lo + m], m)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            whole = per_env(BatchedSupplyNetwork(g_spec, m * world, dev, "int32"), g_init, g_demand, g_actions, m * world)
            shard_check = bool(torch.equal(torch.cat(parts, dim=1), whole))
            assert shard_check, "sharded envs differ from the same envs on one GPU"
        del parts, mine

    env_steps_per_step = n * T * world
    value = env_steps_per_step / (ms_dev / args.steps * 1e-3)
    e2e_value = env_steps_per_step / (ms_e2e / e2e_steps * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = n * T * BYTES_PER_ENV_STEP_ROLLOUT / (kernel_ms * 1e-3) / 1e9

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = single_step_numbers(torch, dev, peak)
    # the other BASELINE configs that shard over the GPUs, under the same launch at every N (collective: all ranks)
    blocks = {}
    h2d = init.nbytes + demand.nbytes + actions.nbytes
    d2h = h_traj_obs.numel() * 4 + h_traj_rew.numel() * 4
    if not args.no_configs:
        del traj_obs, traj_rew, d_actions, h_traj_obs, h_traj_rew
        torch.cuda.empty_cache()
        blocks["config3"] = config3_block(torch, world, rank, dev, 40, 4)
        blocks["config5"] = config5_block(torch, world, rank, dev, 256, 128)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)            # the CPU baseline gets every host core again
            threads = os.cpu_count() or 1
            v, sample = cpu_port_throughput(init, demand, actions, threads, 10.0)
            v1, _ = cpu_port_throughput(init, demand, actions, 1, 3.0)
            cpu = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample,
                   "single_thread_value": v1,
                   "python_reference": None if args.no_python_ref else python_reference_throughput()}
        line = {
            "metric": "env-steps/sec (batched beer game)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(n=n, T=T),
                       "envs_per_gpu": n, "periods": T, "step": "one sn_episode launch (in-kernel reset + 100 periods, 1 episode of the batch)",
                       "l2": "inputs+outputs ~0.9 GB per step >> 126 MB L2, no flush needed",
                       "parallelism": f"env-sharded x{world}, no data-path collective; all-reduce of the 16-double statistics vector every {args.stats_interval} episodes"
                                      + ("" if world == 1 else " (one kernel over peer memory)" if peer_ar is not None else " (NCCL)")},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "api": "BatchedSupplyNetwork.episode_host (pinned host tensors in, trajectory out on the host; H2D/compute/D2H pipelined)", "chunks": args.e2e_chunks,
                    "cpu_binding": (f"{len(cpus)} CPUs next to the GPU (NVML affinity)" if cpus else "none")},
            # our kernels inside the timed region: K episode launches (+ the statistics all-reduce launches at N > 1)
            "gpu_launches": args.steps + ((args.steps // args.stats_interval + 1) if peer_ar is not None else 0),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n), "traffic_source": "static ncu capture (profiles/ncu_traffic.json), not measured in this run",
                         "algorithmic_bytes": n * T * BYTES_PER_ENV_STEP_ROLLOUT, "kernel": "k_rollout<int,ACTIONS,IDENT> via sn_episode", "kernel_ms": kernel_ms,
                         "kernel_ms_how": "CUDA events around the K back-to-back launches of the timed region on this rank / K",
                         "kernel_ms_event_pair_per_launch": kernel_ms_event_pairs,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP_ROLLOUT, "peak_source": peak_src},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        if world > 1:
            line["per_rank_ms_per_step"] = per_rank_ms      # device time of every rank (value uses the max)
            line["stats_exchange"] = stats_exchange
            line["shard_invariance"] = {"ok": shard_check, "envs": 8192 * world,
                                        "what": "every env of a global batch sharded by index over the ranks equals, bit for bit "
                                                "(position-weighted trajectory checksum, episode return), the same env when "
                                                "rank 0 runs the whole batch on one GPU"}
        line.update(extra)
        line.update(blocks)
        print(json.dumps(line))
    if world > 1:
        if peer_ar is not None:
            peer_ar.close()
        dist.destroy_process_group()


def single_step_numbers(torch, dev, peak):
    """Secondary figures: the single-period step kernel (VecEnv mode, 345 B/env-step) at the config-2
    size (L2-resident, launch-bound) and at an HBM-honest 4M envs, plus the in-kernel base-stock
    heuristic rollout (config 3 mode, per GPU)."""
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.simulator import BasePolicyParams, BatchedSupplyNetwork
    out = {}
    for tag, n in (("single_step_65536", 65536), ("single_step_4M", 4 * 1024 * 1024)):
        spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
        sim = BatchedSupplyNetwork(spec, n, dev, "int32")
        g = torch.Generator(device=dev); g.manual_seed(1)
        sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device=dev, dtype=torch.int32))
        sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device=dev, dtype=torch.int32))
        acts = torch.randint(0, 17, (n, 4), generator=g, device=dev, dtype=torch.int32)
        res = {}
        for with_obs in (False, True):
            sim.reset()
            steps = T - 1
            obs_arg = sim.obs if with_obs else None
            import ctypes as C
            from multi_echelon_drl_b200 import _lib

            def one(t):
                _lib.check(sim.lib.sn_step(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, t, C.c_void_p(sim.state.data_ptr()),
                                           C.c_void_p(acts.data_ptr()), C.c_void_p(sim.demand[t + 1].data_ptr()),
                                           C.c_void_p(obs_arg.data_ptr()) if with_obs else None,
                                           C.c_void_p(sim.reward.data_ptr()), None, None, None, None,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            for t in range(5):
                one(t)
            torch.cuda.synchronize()
            # the per-period launches are captured in a CUDA graph: at 65,536 envs one period is a few
            # microseconds of work, far below the cost of issuing a launch from Python
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for t in range(5, steps):
                    one(t)
            graph.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            s.record()
            for _ in range(reps):
                graph.replay()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / (reps * (steps - 5))
            bps = BYTES_PER_ENV_STEP_SINGLE + (112 if with_obs else 0)
            res["with_obs" if with_obs else "state_only"] = {
                "env_steps_per_s": n / (ms * 1e-3), "us_per_launch": ms * 1e3,
                "achieved_gbs": n * bps / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": n * bps / (ms * 1e-3) / 1e9 / peak,
                "bytes_per_env_step": bps, "launch": "CUDA graph of 94 sn_step launches",
                "kernel": ("k_step_tiled (persistent, TMA tensor-map tiles, mbarrier pipeline)" if n * (bps + 0) >= (112 << 20)
                           else "k_step (one thread per env; the batch is L2-resident)")}
        out[tag] = res
        del sim
    # the same period step at 4 Mi envs on other networks: chain with other lead times (57 state
    # words), two-level distribution tree (64 words, two demand rows), general kernels, observation out
    from dataclasses import replace as _replace
    from multi_echelon_drl_b200.spec import tree_spec as _tree_spec
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    n = 4 * 1024 * 1024
    base = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
    nets = {"other_leadtimes": (_replace(base, info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1)), 57),
            "distribution_tree": (_tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20),
                                              dict(name="dist_b", demand=(0, 5), target=14), dict(name="maker", target=25)],
                                             [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2),
                                              ("dist_a", "shop", 2, 2)], ["shop", "dist_a", "dist_b", "maker"], True), 64)}
    out["single_step_4M_other_networks"] = {}
    os.environ["SN_NO_SPECIALIZED"] = "1"
    for name, (sp, words) in nets.items():
        sim = BatchedSupplyNetwork(sp, n, dev, "int32")
        g = torch.Generator(device=dev); g.manual_seed(1)
        sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device=dev, dtype=torch.int32))
        sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device=dev, dtype=torch.int32))
        acts = torch.randint(0, 17, (n, 4), generator=g, device=dev, dtype=torch.int32)
        sim.reset()

        def one(t):
            _lib.check(sim.lib.sn_step(C.byref(sim._c_spec), sim.dtype_code, n, sim.ld, t, C.c_void_p(sim.state.data_ptr()),
                                       C.c_void_p(acts.data_ptr()), C.c_void_p(sim.demand[t + 1].data_ptr()),
                                       C.c_void_p(sim.obs.data_ptr()), C.c_void_p(sim.reward.data_ptr()), None, None, None, None,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for t in range(3):
            one(t)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for t in range(3, 43):
                one(t)
        graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 120
        bps = 2 * words * 4 + 16 + 4 * sp.n_sources + 4 + 112
        out["single_step_4M_other_networks"][name] = {
            "us_per_launch": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3), "bytes_per_env_step": bps,
            "frac_of_hbm_peak": n * bps / (ms * 1e-3) / 1e9 / peak, "kernel": "k_step_tiled (same pipeline, wider state box)"}
        del sim
    os.environ.pop("SN_NO_SPECIALIZED", None)
    # the literal drop-in call: SB3-style VecEnv.step with numpy in / numpy out (pinned staging, H2D of the
    # actions and D2H of obs + reward every step), BeerGameUniformMultiFacilityFullInfo-v0, 65,536 envs
    from multi_echelon_drl_b200.register_envs import make
    venv = make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=65536, dtype="float32", seed=0, output="numpy")
    venv.reset()
    a_np = np.full((65536, 4), 4.0, dtype=np.float32)
    for _ in range(5):
        venv.step(a_np)
    t0 = time.perf_counter()
    reps = 190
    for _ in range(reps):
        venv.step(a_np)
    dt = time.perf_counter() - t0
    out["vecenv_numpy_step_65536"] = {"env_steps_per_s": 65536 * reps / dt, "us_per_step": dt / reps * 1e6,
                                      "h2d_bytes_per_step": a_np.nbytes, "d2h_bytes_per_step": 65536 * (28 * 4 + 4),
                                      "note": "wall clock incl. host staging copies, auto-reset every 100 steps"}
    del venv
    # config-3 mode: heuristic rollout, stats only (ALU/latency bound, ~4 B/env-step of HBM traffic)
    n = 131072
    spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
    sim = BatchedSupplyNetwork(spec, n, dev, "int32")
    g = torch.Generator(device=dev); g.manual_seed(2)
    sim.demand.copy_(torch.randint(0, 9, sim.demand.shape, generator=g, device=dev, dtype=torch.int32))
    sim.init.copy_(torch.randint(0, 9, sim.init.shape, generator=g, device=dev, dtype=torch.int32))
    pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
    for _ in range(3):
        sim.episode(T, policy=pol, write_state=False)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    reps = 20
    for _ in range(reps):
        sim.episode(T, policy=pol, write_state=False)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    out["heuristic_rollout_131072"] = {"env_steps_per_s": n * T / (ms * 1e-3), "ms_per_episode": ms,
                                       "bound": "alu/latency (state on-chip, 4 B/env-step HBM)",
                                       "config": "config 3 per-GPU share: 131,072 envs, base-stock [19,20,20,14] in-kernel"}
    del sim
    # config 3 end to end on one GPU's share: 10 episodes, fresh device-generated demand / initial state per episode
    from multi_echelon_drl_b200.distributed import summarize
    venv = make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=131072, dtype="int32", seed=3)
    venv.run_policy_episodes(pol, 2)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    st = venv.run_policy_episodes(pol, 10)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    summ = summarize(st)
    out["heuristic_10_episodes_131072"] = {"env_steps_per_s": 131072 * T * 10 / (ms * 1e-3), "ms_total": ms,
                                           "mean_episode_return": summ["mean_episode_return"],
                                           "std_episode_return": summ["std_episode_return"],
                                           "note": "incl. per-episode demand/init generation (torch Philox) and episode moments"}
    del venv
    # config-4 mode: 262,144 envs x 2 items (independent item chains, per-item costs), single-step API; the per-env
    # reward is reduced over the two adjacent item chains inside the step kernel
    from multi_echelon_drl_b200.multi_item import MultiItemSupplyNetwork
    from dataclasses import replace
    n = 262144
    specs = [spec, replace(spec, holding_cost=(0.75,) * 4)]
    # per chain: state 160 r + 160 w, action 16, demand 4, chain reward 4, observation 112, episode return 8 r + 8 w,
    # env reward 4 per env = 2 per chain
    bpc = BYTES_PER_ENV_STEP_SINGLE + 112 + 16 + 2
    res = {}
    for mode in ("eager", "graph"):
        mi = MultiItemSupplyNetwork(specs, n, dev, "int32", use_graph=(mode == "graph"))
        mi.sim.demand.random_(0, 9, generator=g)
        mi.sim.init.random_(0, 9, generator=g)
        mi.action_buffer.copy_(torch.randint(0, 17, (n, 8), generator=g, device=dev, dtype=torch.int32))
        acts = mi.action_buffer
        mi.reset()
        for _ in range(5):
            mi.step(acts)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        reps = 60
        for _ in range(reps):
            mi.step(acts)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        res[mode] = {"us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3), "achieved_gbs": 2 * n * bpc / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": 2 * n * bpc / (ms * 1e-3) / 1e9 / peak}
        if mode == "graph":
            # the same launch 90 times in ONE graph (how single_step_* above are timed): no Python between the periods
            sim = mi.sim
            mi.reset()
            sim._graph_period_valid = True
            sim.ctl[0:1].fill_(sim.period)
            torch.cuda.synchronize()
            many, periods = torch.cuda.CUDAGraph(), 90
            with torch.cuda.graph(many):
                for _ in range(periods):
                    sim.launch_step_device_period(mi._act, sim.obs, sim.reward, env_reward=mi.reward)
            best = 1e9
            for _ in range(3):
                mi.reset()
                sim.ctl[0:1].fill_(0)
                torch.cuda.synchronize()
                s.record()
                many.replay()
                e.record()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e) / periods)
            assert sim.device_flags() == 0
            res["graph_of_periods"] = {"us_per_step": best * 1e3, "env_steps_per_s": n / (best * 1e-3),
                                       "achieved_gbs": 2 * n * bpc / (best * 1e-3) / 1e9,
                                       "frac_of_hbm_peak": 2 * n * bpc / (best * 1e-3) / 1e9 / peak,
                                       "launch": "CUDA graph of 90 period launches (device period counter)"}
            sim._graph_period_valid = False
            del many
        del mi
    g90 = res["graph_of_periods"]
    out["multi_item_262144x2_step"] = {
        "env_steps_per_s": g90["env_steps_per_s"], "item_chain_steps_per_s": 2 * g90["env_steps_per_s"],
        "us_per_step": g90["us_per_step"], "frac_of_hbm_peak": g90["frac_of_hbm_peak"],
        "achieved_gbs": g90["achieved_gbs"], "bytes_per_chain_step": bpc, "launch": g90["launch"],
        "python_issued_per_period": res["eager"], "graph_replay_per_period": res["graph"],
        "config": "config 4: 262,144 envs x 2 items (one interleaved batch of 524,288 chains, per-item costs and the per-env "
                  "reward sum in-kernel), obs 56 / action 8, k_step_tiled; timed like single_step_*: 90 period launches in one "
                  "CUDA graph (device period counter); python_issued_per_period = MultiItemSupplyNetwork.step called per "
                  "period from Python, graph_replay_per_period = one replayed single-period graph per call"}
    # other networks, same fixed-action rollout with trajectory out at config-2 size
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    n = N_ENVS
    tree = tree_spec([dict(name="shop", demand=(0, 8), target=19), dict(name="dist_a", target=20),
                      dict(name="dist_b", demand=(0, 5), target=14), dict(name="maker", target=25)],
                     [("ext", "maker", 1, 2), ("maker", "dist_b", 1, 2), ("maker", "dist_a", 2, 2), ("dist_a", "shop", 2, 2)],
                     ["shop", "dist_a", "dist_b", "maker"], True)
    big = tree_spec([dict(name="hub", target=60), dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=14),
                     dict(name="r3", demand=(0, 4), target=10), dict(name="r4", demand=(0, 3), target=8),
                     dict(name="r5", demand=(0, 2), target=6), dict(name="factory", target=40),
                     dict(name="annex", demand=(0, 5), target=12)],
                    [("ext", "factory", 1, 2), ("factory", "annex", 1, 3), ("factory", "hub", 2, 2), ("hub", "r3", 2, 1),
                     ("hub", "r1", 2, 2), ("hub", "r5", 3, 1), ("hub", "r2", 1, 1), ("hub", "r4", 1, 2)],
                    ["hub", "r1", "r2", "r3", "r4", "r5", "factory", "annex"], True)
    five = tree_spec([dict(name="hub", target=60), dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=14),
                      dict(name="r3", demand=(0, 4), target=10), dict(name="factory", target=40)],
                     [("ext", "factory", 1, 2), ("factory", "hub", 2, 2), ("hub", "r3", 2, 1), ("hub", "r1", 2, 2), ("hub", "r2", 1, 1)],
                     ["hub", "r1", "r2", "r3", "factory"], True)
    others = {"rollout_float32_65536": spec,
              "rollout_other_leadtimes_65536": replace(spec, info_leadtime=(3, 2, 1, 2), ship_leadtime=(2, 3, 4, 1)),
              "rollout_distribution_tree_65536": tree, "rollout_distribution_tree_65536_specialised": tree,
              "rollout_tree_5_nodes_65536": five,
              "rollout_tree_8_nodes_65536": big, "rollout_tree_8_nodes_65536_specialised": big}
    t_rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    from multi_echelon_drl_b200 import specialize
    for key, sp in others.items():
        qdtype = "float32" if "float32" in key else "int32"
        acts = torch.randint(0, 17, (T, n, sp.n_agents), generator=g, device=dev, dtype=torch.int32)
        if qdtype == "float32":                     # TD3's quantities: fractional orders
            acts = acts.to(torch.float32) + torch.rand(acts.shape, generator=g, device=dev)
        t_obs = torch.empty((T, n, sp.obs_dim), dtype=torch.float32, device=dev)
        # the general tree kernels take the topology at run time; a library compiled for this very network
        # (multi_echelon_drl_b200/specialize.py, prebuilt by __graft_entry__.build()) is measured beside them
        os.environ["SN_NO_SPECIALIZED"] = "0" if key.endswith("_specialised") else "1"
        specialize._loaded.clear()
        sim = BatchedSupplyNetwork(sp, n, dev, qdtype)
        os.environ.pop("SN_NO_SPECIALIZED", None)
        if key.endswith("_specialised") and sim.fast is None:
            out[key] = {"unavailable": "no specialised library for this network has been built"}
            continue
        sim.load_episode(*bulk_episode_inputs(sp, n, np.random.RandomState(0)))
        for _ in range(3):
            sim.episode(T, actions=acts, traj_obs=t_obs, traj_reward=t_rew, write_state=False)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(10):
            sim.episode(T, actions=acts, traj_obs=t_obs, traj_reward=t_rew, write_state=False)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        nbytes = n * T * (4 * sp.n_agents + 4 * sp.n_sources + 4 * sp.obs_dim + 4)
        out[key] = {"env_steps_per_s": n * T / (ms * 1e-3), "kernel_ms": ms, "hbm_gbs": nbytes / (ms * 1e-3) / 1e9,
                    "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                    "bound": ("hbm (the headline kernel with float32 quantities)" if "float32" in key else
                              "hbm / issue (topology and lead times folded at compile time)" if key.endswith("_specialised") else
                              "instruction issue (run-time topology / lead times as 0/1 multipliers)"),
                    "config": ("config-2 rollout with float32 quantities and fractional orders (what TD3 feeds the env)" if "float32" in key else
                               "5-node tree: factory -> hub -> 3 retailers (one lane per node, five lanes per env: six envs per warp)" if "5_nodes" in key else
                               "serial chain, lead times 3,2,1,2 / 2,3,4,1 (generic kernels)" if "leadtimes" in key else
                               "two-level distribution tree, two demand sources (library specialised for this network)" if key == "rollout_distribution_tree_65536_specialised" else
                               "8-node tree: factory -> {annex, hub -> 5 retailers} (library specialised for this network: eight "
                               "facilities per thread, register-resident)" if key.endswith("_specialised") else
                               "two-level distribution tree, two demand sources (general tree kernels)" if "distribution" in key else
                               "8-node tree: factory -> {annex, hub -> 5 retailers}, 6 demand sources, obs 56 / action 8 "
                               "(one lane per node)")}
        del sim
    return out


def config3_block(torch, world, rank, dev, steps, warmup):
    """BASELINE config 3: base-stock (HGE heuristic) rollouts, 131,072 envs per GPU (1,048,576 on 8), policy evaluated
    in-kernel, fresh device-generated demand / initial state per episode, statistics all-reduced over NCCL at the end.
    One step = one 100-period episode of every env.  Collective over the ranks: every rank calls it, every rank gets
    the result (time = max over ranks)."""
    import torch.distributed as dist
    from multi_echelon_drl_b200.distributed import allreduce_stats, summarize
    from multi_echelon_drl_b200.register_envs import make
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    n = 131072
    venv = make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="int32", seed=1000 + rank)
    pol = BasePolicyParams([19, 20, 20, 14], 0, 16, "a")
    venv.run_policy_episodes(pol, max(warmup, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    stats = venv.run_policy_episodes(pol, steps)
    allreduce_stats(stats)
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    summ = summarize(stats)
    assert summ["env_steps"] == n * T * steps * world
    return {"value": n * T * steps * world / (ms.item() * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": steps,
            "envs_total": n * world, "ms_per_step": ms.item() / steps,
            "mean_episode_return": summ["mean_episode_return"], "std_episode_return": summ["std_episode_return"],
            "qty_range_flags": float(stats[13].item()),
            "workload": f"config3: base-stock [19,20,20,14] rollouts, {n} envs/GPU x {world} GPUs = {n * world} envs, policy "
                        "in-kernel, demand / initial state regenerated on the device per episode; one step = one 100-period "
                        "episode of every env (generation + 1 sn_episode launch)",
            "bound": "alu/issue (state on-chip, 4 B/env-step of HBM)"}


def config5_block(torch, world, rank, dev, steps, warmup):
    """BASELINE config 5: TD3 with heuristic-guided exploration on 4,096 complex-scenario envs per GPU (hyper-parameters
    of setup_example_centralized.yaml), env -> VecNormalize -> policy -> replay buffer -> update all on the device.
    One step = one vector step of every env plus one gradient update (train_freq 32, 32 updates per 32 steps).
    Collective over the ranks (gradients are averaged over NCCL inside the update graph)."""
    from multi_echelon_drl_b200.register_envs import make
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 4096
    env = VecNormalize(make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=1000 * rank,
                            use_graph=True), clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(batch_size=128, net_arch=(768,), learning_rate=1e-3, tau=0.01, train_freq=32, gradient_steps=32,
                    action_noise_std=0.4, gamma=0.95, hge_rate_at_start=0.5, seed=rank)
    model = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), cfg, ddp=world > 1)
    steps = max(32, steps // 32 * 32)
    warm = max(128, warmup // 32 * 32)                    # covers the eager warm-up cycles and the graph captures
    model.learn(n * warm)
    torch.cuda.synchronize()
    model.timers = type(model.timers)()
    t0 = time.perf_counter()
    model.learn(model.num_timesteps + n * steps)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    split = model.time_split()
    # device time of the env step alone: the env's captured graph (step + VecNormalize in one cooperative kernel)
    # replayed back to back inside one episode; the in-loop figure above it includes the Python launch path
    venv = env.venv
    env.reset()
    for _ in range(5):
        env.step(venv.action_buffer)
    reps = min(80, venv.sim.T - venv.sim.period - 2)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        venv._graph.replay()
        venv.sim.advance_host_period()
    e.record()
    torch.cuda.synchronize()
    env_step_graph_us = s.elapsed_time(e) / reps * 1e3
    # the same launch 40 times in ONE graph (no launch gap between replays): the kernel's own device time
    sim = venv.sim
    env.reset()
    torch.cuda.synchronize()
    many = torch.cuda.CUDAGraph()
    with torch.cuda.graph(many):
        for _ in range(40):
            sim.launch_step_device_period(venv.action_buffer, venv._obs, sim.reward, vecnorm=env.fused_args(venv._graph_key))
    many.replay()                                         # (the first replay of a new graph pays its upload)
    env_step_kernel_us = 1e9
    for _ in range(3):
        env.reset()
        torch.cuda.synchronize()
        s.record()
        many.replay()
        many.replay()
        e.record()
        torch.cuda.synchronize()
        env_step_kernel_us = min(env_step_kernel_us, s.elapsed_time(e) / 80 * 1e3)
    flags = venv.sim.device_flags()
    del many
    model.close()
    if world > 1:
        w = torch.tensor([wall], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(w, op=torch.distributed.ReduceOp.MAX)
        wall = float(w.item())
    return {"value": n * steps * world / wall, "unit": "env-steps/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": wall / steps * 1e3, "updates": steps,
            "time_split_s": {k: v for k, v in split.items() if k.endswith("_s")},
            "us_per_vector_step": {k[:-2]: round(v / steps * 1e6, 2) for k, v in split.items() if k.endswith("_s")},
            "env_step_graph_gpu_us": round(env_step_graph_us, 2), "env_step_kernel_gpu_us": round(env_step_kernel_us, 2),
            "device_flags": flags,
            "workload": f"config5: TD3+HGE, {n} envs/GPU x {world}, batch 128, one 768-unit hidden layer, train_freq 32 / 32 "
                        "updates, device replay buffer; env step + VecNormalize = ONE replayed cooperative kernel (device "
                        "period counter); per vector step three graph replays (action, env step, store) + one replayed update (fused Adam); one step = one vector step of all envs + "
                        "one gradient update (wall clock, max over ranks); GPU-time split per vector step in us"}


def _init_dist(torch):
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def run_config3(args):
    import torch
    world, rank, dev = _init_dist(torch)
    b = config3_block(torch, world, rank, dev, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec (batched beer game)", "value": b["value"], "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": b["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": b["workload"]}, "gpu_launches": args.steps,
            "mean_episode_return": b["mean_episode_return"], "std_episode_return": b["std_episode_return"],
            "roofline": {"bound": "alu/issue", "note": "state on-chip, 4 B/env-step of HBM: not bandwidth bound; "
                                                       "197 instructions per warp-period, issue slots 66 % busy (profiles/r1_notes.md)"}}))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_config5(args):
    import torch
    world, rank, dev = _init_dist(torch)
    b = config5_block(torch, world, rank, dev, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec (TD3+HGE training loop)", "value": b["value"], "unit": "env-steps/s",
            "n_gpus": world, "steps": b["steps"], "warmup": b["warmup"], "ms_per_step": b["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": b["workload"]}, "time_split_s": b["time_split_s"],
            "us_per_vector_step": b["us_per_vector_step"], "updates": b["updates"],
            "env_step_graph_gpu_us": b["env_step_graph_gpu_us"], "env_step_kernel_gpu_us": b["env_step_kernel_gpu_us"],
            "device_flags": b["device_flags"]}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS, help="envs per GPU (default: BASELINE config 2)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config5"],
                    help="config2 (default, the headline line), config3 (heuristic rollouts, 131,072 envs/GPU) or "
                         "config5 (TD3+HGE training loop, 4,096 envs/GPU)")
    ap.add_argument("--stats-interval", type=int, default=10, help="episodes between statistics all-reduces (N>1)")
    ap.add_argument("--e2e-chunks", type=int, default=5, help="period spans of the pipelined host path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-python-ref", action="store_true", help="skip the reference's own Python loop inside cpu_baseline")
    ap.add_argument("--no-configs", action="store_true", help="skip the config-3 / config-5 blocks of the default line")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's socket")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary single-step / heuristic figures")
    ap.add_argument("--nccl-stats", action="store_true", help="N>1: carry the statistics vector with NCCL instead of the peer-memory kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config3":
        run_config3(args)
    elif args.workload == "config5":
        run_config5(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
