"""Pure-copy ceiling of the host path at 1/2/4/8 GPUs: every rank moves what one config-2 e2e step moves (760 MB
device->host trajectory + 137 MB host->device inputs, pinned memory, two streams, plain cudaMemcpyAsync through
torch's copy_, NO kernels) at the same time as the others.  Run under torchrun; rank 0 prints one JSON line.
If the aggregate here equals what bench.py's e2e arm reaches at the same N, the e2e number is the platform's host-side
ceiling (PCIe root / VM memory path), not the pipeline's.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/probe/pcie_multi_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d2h_bytes, h2d_bytes = 760217600, 136839168          # bench.py: e2e.d2h_bytes_per_step / h2d_bytes_per_step
d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def step(do_d2h=True, do_h2d=True):
    main = torch.cuda.current_stream()
    s1.wait_stream(main)
    s2.wait_stream(main)
    if do_d2h:
        with torch.cuda.stream(s1):
            h_out.copy_(d_out, non_blocking=True)
    if do_h2d:
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)
    main.wait_stream(s1)                      # both directions are in flight together
    main.wait_stream(s2)


def timed(reps=6, **kw):
    step(**kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        step(**kw)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    allms = [torch.zeros_like(ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(allms, ms)
    else:
        allms = [ms]
    return [round(float(x.item()), 3) for x in allms]


res = {"n_gpus": world, "d2h_bytes": d2h_bytes, "h2d_bytes": h2d_bytes}
for tag, kw, nbytes in (("d2h_only", dict(do_h2d=False), d2h_bytes), ("h2d_only", dict(do_d2h=False), h2d_bytes),
                        ("both", {}, d2h_bytes + h2d_bytes)):
    per_rank = timed(**kw)
    slowest = max(per_rank)
    res[tag] = {"per_rank_ms": per_rank, "slowest_ms": slowest, "aggregate_gbs": round(world * nbytes / slowest / 1e6, 1),
                "per_gpu_gbs": round(nbytes / slowest / 1e6, 1)}
# the e2e arm's unit: one step = 6,553,600 env-steps per rank
res["env_steps_per_s_if_copies_were_all"] = round(world * 65536 * 100 / (res["both"]["slowest_ms"] * 1e-3))
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
