"""Concurrent pinned host<->device copy bandwidth of the box, per rank and aggregate (torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/host_copy_bw.py

Answers what bounds the 8-GPU end-to-end number (bench.py `e2e`): every rank moves ~0.33 MB of actions to and ~0.86 MB
of results from its GPU per step through ONE host memory system. Measured: each rank alone, then all ranks at once,
for large copies (64 MiB, bandwidth) and for the bench's own transfer sizes (latency + bandwidth), both directions
separately and together. Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    sys.stdout.flush()
    os.dup2(saved, 1)
if os.environ.get("FLOCK_PIN_CORES", "1") != "0" and world > 1:
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // world)
    os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)


def barrier():
    if world > 1:
        dist.barrier()


def run(nbytes_h2d, nbytes_d2h, iters, both=True, h2d=True, d2h=True):
    hs = torch.empty(max(nbytes_h2d, 1), dtype=torch.uint8).pin_memory()
    hd = torch.empty(max(nbytes_d2h, 1), dtype=torch.uint8).pin_memory()
    ds = torch.empty(max(nbytes_h2d, 1), dtype=torch.uint8, device=dev)
    dd = torch.empty(max(nbytes_d2h, 1), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(3):
        ds.copy_(hs, non_blocking=True)
        hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(s_in):
                ds.copy_(hs, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    barrier()
    moved = iters * ((nbytes_h2d if h2d else 0) + (nbytes_d2h if d2h else 0))
    return moved / dt / 1e9, dt / iters * 1e6


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


res = {"world": world, "host_threads": len(os.sched_getaffinity(0)), "os_cpu_count": os.cpu_count()}
BIG = 64 << 20
STEP_IN, STEP_OUT = 327680, 864256            # cfg2: actions in, obs | reward | dones out per step
for label, a, b, it, kw in (("big_h2d", BIG, BIG, 40, dict(d2h=False)), ("big_d2h", BIG, BIG, 40, dict(h2d=False)),
                            ("big_both", BIG, BIG, 40, {}), ("step_sized_both", STEP_IN, STEP_OUT, 2000, {})):
    # all ranks at once
    gbs, us = run(a, b, it, **kw)
    allg = gather(gbs)
    res[label + "_concurrent"] = {"per_rank_GBps": allg, "aggregate_GBps": sum(allg), "us_per_iter_rank0": us}
    # one rank at a time (rank 0 only is timed alone; the others idle at the barrier inside run())
    if world > 1:
        if rank == 0:
            hs = None
        alone = []
        for r in range(min(world, 2)):          # ranks 0 and 1 alone: enough to see the single-link figure
            if rank == r:
                # temporarily behave as a single-rank job
                w_save = world
                globals()["world"] = 1
                g1, _ = run(a, b, it, **kw)
                globals()["world"] = w_save
            else:
                g1 = 0.0
            dist.barrier()
            alone.append(max(gather(g1)))
        res[label + "_alone_GBps"] = alone
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
