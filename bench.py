#!/usr/bin/env python
"""Throughput benchmark of the flocking-env hot path (agent-steps/s), see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...        # CPU arm: the C oracle on the host cores

A "step" is ONE fused-kernel pass of `step()` over one batch of E envs x N agents (random actions
already resident in HBM). To keep the working set larger than the 126 MB L2, the bench cycles over
a ring of R independent env batches (step s touches batch s % R), all on one stream, replayed from
CUDA graphs so that Python launch overhead is not what is measured. Under torchrun every rank owns
its own ring on its own GPU (envs shard with no data-path collective; the only collective is the
NCCL all-reduce of the episode statistics after the timed region) -> weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs: density-scaled synthetic worlds, random actions
WORKLOADS = {
    # name: variant, E per GPU, N, k, collision, range_start, sensor, extra VecEnv kwargs, bytes/agent-step
    "cfg2": dict(variant="v2", E=4096, N=10, k=4, cd=2.5, rs=(0, 50), sr=14.0, kw={}, bytes=53,
                 desc="gym_flock_v2 batched 4096 envs x 10 agents, k=4, random actions (BASELINE configs[1])"),
    "cfg3": dict(variant="uw", E=4096, N=32, k=3, cd=0.5, rs=(0, 200), sr=7.0, kw={}, bytes=125,
                 env_kw=dict(track_neighbors=False),   # the reference discards the indices in uw (gym_flock_uw.py:141-144)
                 desc="gym_flock_uw 4096 envs x 32 agents, k=3, random actions, materialised (N,4,k) window"),
    "cfg4": dict(variant="uwd", E=8192, N=16, k=4, cd=0.5, rs=(0, 100), sr=7.0,
                 kw=dict(reset_collision_distance=1.0), bytes=49,
                 env_kw=dict(track_neighbors=False),   # ... and in uw_discrete (gym_flock_uw_discrete.py:189-192)
                 desc="gym_flock_uw_discrete 8192 envs x 16 agents, k=4, random action ids, Philox actuation noise"),
    "cfg5": dict(variant="v2", E=64, N=2048, k=8, cd=0.05, rs=(0, 2000), sr=100.0, kw={}, bytes=69,
                 desc="gym_flock_v2 large swarm 64 envs x 2048 agents, k=8 (tiled path, exact box-pruned k-NN)"),
}
L2_BYTES = 126 * 1024 * 1024
DT = 0.1


def _bf16_peak():
    """dense bf16 TFLOP/s: measured cuBLAS burst figure of MEASURED_PEAKS.json, else the 2250 nominal"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"])
    except Exception:
        return 2250.0


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.max_mhz = period_s, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ---------------------------------------------------------------------------------------------------
# CPU arm (oracle/flock_oracle.c = a C port of the reference step, all host threads)
# ---------------------------------------------------------------------------------------------------
def cpu_run(w, steps, warmup, budget_s, nthreads=None):
    """Time `steps` oracle steps on a bounded sample of the workload's envs; returns a dict."""
    from oracle import flock_oracle as fo

    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle sets the
    # OpenMP team size explicitly, so that default does not apply)
    threads = nthreads or len(os.sched_getaffinity(0)) or fo.max_threads()
    E_full = w["E"]
    # calibrate the sample size so that warmup+steps fit the time budget
    probe_E = min(E_full, max(threads * 4, 64))
    env = fo.OracleEnv(w["variant"], probe_E, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"],
                       seed=0x5EED, nthreads=threads, **w["kw"])
    env.reset()
    a = env.random_actions()
    t0 = time.perf_counter()
    for _ in range(3):
        env.step(a, DT)
    per_env_step = (time.perf_counter() - t0) / 3 / probe_E
    E = int(max(threads, min(E_full, budget_s / max(per_env_step * (steps + warmup), 1e-12))))
    env = fo.OracleEnv(w["variant"], E, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"],
                       seed=0x5EED, nthreads=threads, **w["kw"])
    env.reset()
    acts = [env.random_actions(i) for i in range(8)]
    for i in range(warmup):
        env.step(acts[i % 8], DT)
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(acts[i % 8], DT)
    dt = time.perf_counter() - t0
    return dict(value=E * w["N"] * steps / dt, seconds=dt, envs=E, threads=threads,
                sample=f"{steps} steps of {E} of the {E_full} envs x {w['N']} agents (oracle/flock_oracle.c, OpenMP over envs)")


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 2000), min(args.warmup, 50)
    r = cpu_run(w, steps, warmup, budget_s=60.0)
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": r["value"], "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["seconds"] / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "envs_per_step": r["envs"], "agents": w["N"], "k": w["k"]},
        "cpu_baseline": {"value": r["value"], "unit": "agent-steps/s", "cores": r["threads"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference itself is Python/PyTorch and cannot travel to the GPU box; this arm times the C port "
                "of its step (oracle/), which is FASTER than the reference's own dispatch-bound step (BASELINE.md s3)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def pairs_evaluated_fraction(w, E, device, steps=32):
    """Share of the N*(N-1) pairs per row the sensing kernel evaluates (1.0 for the all-pairs kernels),
    counted by the kernel itself on a separate env (the counter costs an atomic per warp)."""
    import torch
    from marl_range_flocking_b200 import VecEnv

    env = VecEnv(w["variant"], E, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=0xC0,
                 device=device, **w["kw"], **w.get("env_kw", {}))
    env.reset()
    env.pairs_evaluated()                       # first query switches the counter on
    acts = [env.random_actions(i) for i in range(2)]
    for t in range(steps):
        env.step(acts[t & 1], DT)
    torch.cuda.synchronize(device)
    pairs = env.pairs_evaluated()
    if pairs == 0:                              # a kernel without pruning evaluates everything
        return 1.0
    return pairs / (steps * E * w["N"] * (w["N"] - 1))


def build_ring(w, E, ring, device, env_offset, seed=0x5EED):
    import torch
    from marl_range_flocking_b200 import VecEnv

    envs, acts = [], []
    for r in range(ring):
        env = VecEnv(w["variant"], E, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"],
                     seed=seed + r, env_offset=env_offset, device=device, **w["kw"], **w.get("env_kw", {}))
        env.reset()
        envs.append(env)
        acts.append([env.random_actions(i) for i in range(2)])
    torch.cuda.synchronize(device)
    return envs, acts


def capture(envs, acts, n, start):
    """CUDA graph of n consecutive bench steps (step s -> batch s % R)."""
    import torch

    g = torch.cuda.CUDAGraph()
    R = len(envs)
    before = sum(e.launch_count for e in envs)
    with torch.cuda.graph(g):
        for s in range(start, start + n):
            envs[s % R].step(acts[s % R][(s // R) & 1], DT)
    return g, sum(e.launch_count for e in envs) - before       # kernel nodes in the graph


def timed_graph_steps(envs, acts, steps, warmup, device, dist_barrier):
    import torch

    chunk = min(steps, 1024)
    full, rem = divmod(steps, chunk)
    g_full, n_full = capture(envs, acts, chunk, 0)
    g_rem, n_rem = capture(envs, acts, rem, 0) if rem else (None, 0)
    wfull, wrem = divmod(warmup, chunk)
    g_w, _ = capture(envs, acts, wrem, 0) if wrem else (None, 0)
    for _ in range(wfull):
        g_full.replay()
    if g_w is not None:
        g_w.replay()
    torch.cuda.synchronize(device)
    dist_barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    ev0.record()
    for _ in range(full):
        g_full.replay()
    if g_rem is not None:
        g_rem.replay()
    ev1.record()
    torch.cuda.synchronize(device)
    dist_barrier()
    # graph replays do not go through flock_step again: launches = kernel nodes of the replayed graphs
    # (one fused kernel per step for N <= 32; integrate pre-pass + sensing kernel, and a row-order
    # refresh every 16 steps of an env, for larger swarms)
    return ev0.elapsed_time(ev1), steps, full * n_full + n_rem


def timed_e2e(env, w, steps, warmup, device):
    """Host buffers in / out through flock_step_host (C ABI): H2D actions, fused step, D2H results."""
    import torch

    E, N = env.num_envs, env.num_particles
    host_acts = [env.random_actions(i).cpu().pin_memory() for i in range(4)]
    torch.cuda.synchronize(device)
    for i in range(warmup):
        env.step_host(host_acts[i % 4], DT)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        env.step_host(host_acts[i % 4], DT)
    ev1.record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t0
    h2d = host_acts[0].numel() * 4
    d2h = env._obs.numel() * 4 + E * N * 4 + E * N + E
    return max(ev0.elapsed_time(ev1) * 1e-3, wall), h2d, d2h


def timed_e2e_pipelined(pipe, steps, warmup, device):
    """len(pipe) env batches in flight (VecEnv.step_host_async / wait_host -> flock_step_host_async): every
    step still copies its actions from pinned host memory and reads its results back to the host, but one
    batch's result transfer overlaps the other batches' action transfer and fused step. Returns seconds for
    `steps` steps (steps counts single-batch steps, like timed_e2e)."""
    import torch

    acts = [pipe[0].random_actions(i).cpu().pin_memory() for i in range(4)]
    torch.cuda.synchronize(device)
    D = len(pipe)
    for e in pipe:
        e.step_host_async(acts[0], DT)
    for i in range(warmup):
        e = pipe[i % D]
        e.wait_host()
        e.step_host_async(acts[i % 4], DT)
    for e in pipe:
        e.wait_host()
    torch.cuda.synchronize(device)
    sink = 0
    t0 = time.perf_counter()
    for e in pipe:
        e.step_host_async(acts[0], DT)
    for i in range(steps - D):
        e = pipe[i % D]
        _, _, (_, env_done), _ = e.wait_host()
        sink += int(env_done[0])                 # the host really looks at the result before stepping on
        e.step_host_async(acts[i % 4], DT)
    for e in pipe:
        e.wait_host()
    wall = time.perf_counter() - t0
    torch.cuda.synchronize(device)
    return wall


def run_gpu(args, w):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at communicator creation; stdout carries ONE JSON line, so the
        # file descriptor points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

        def barrier():
            dist.barrier()
    else:
        def barrier():
            return None

    E, N = w["E"], w["N"]
    bytes_per_batch = E * N * (w["bytes"] + 40)           # + nn_idx, velocities, episode counters
    ring = args.ring or max(2, -(-int(1.25 * L2_BYTES) // bytes_per_batch))
    envs, acts = build_ring(w, E, ring, device, env_offset=rank * E)

    sampler = ClockSampler(local if os.environ.get("CUDA_VISIBLE_DEVICES") is None else 0)
    # map the torch device to its NVML index through the UUID when possible
    try:
        import pynvml
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        pynvml.nvmlInit()
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(h)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid in u:
                sampler.h = h
                sampler.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        pass
    sampler.start()
    ms, steps, launches = timed_graph_steps(envs, acts, args.steps, args.warmup, device, barrier)
    clocks = sampler.stop()

    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    agent_steps = world * E * N * steps
    value = agent_steps / (ms_max * 1e-3)

    # end-to-end through the host-buffer C ABI call (every rank, max time)
    e2e_steps = min(args.steps, 2000)
    e2e_s, h2d, d2h = timed_e2e(envs[0], w, e2e_steps, min(args.warmup, 20), device)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_value = world * E * N * e2e_steps / float(t.item())
    e2e_value, e2e_api = e2e_sync_value, "VecEnv.step_host -> flock_step_host (pinned host buffers, sync per step)"
    if len(envs) >= 2:
        depth = min(len(envs), max(2, int(os.environ.get("FLOCK_E2E_DEPTH", "4"))))
        t = torch.tensor([timed_e2e_pipelined(envs[:depth], e2e_steps, min(args.warmup, 20), device)],
                         dtype=torch.float64, device=device)
        if use_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pipe_value = world * E * N * e2e_steps / float(t.item())
        if e2e_pipe_value > e2e_sync_value:
            e2e_value = e2e_pipe_value
            e2e_api = (f"VecEnv.step_host_async / wait_host -> flock_step_host_async: {depth} env batches in flight, every step "
                       "moves its actions from pinned host memory and its results back to pinned host memory, wall clock")
    else:
        e2e_pipe_value = None

    # the one collective of the system: all-reduce of the episode statistics (NCCL over NVLink)
    stats = envs[0].stats_tensor().clone()
    if use_dist:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)

    extra = {}
    # what a real rollout sees: ONE env batch stepped over and over, its state resident in L2 (informational: the
    # headline above cycles over a ring larger than L2, as the timing rules require)
    if rank == 0:
        g1, _ = capture(envs[:1], acts[:1], 512, 0)
        g1.replay()
        torch.cuda.synchronize(device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(4):
            g1.replay()
        ev1.record()
        torch.cuda.synchronize(device)
        t1 = ev0.elapsed_time(ev1) * 1e-3 / (4 * 512)
        extra["l2_resident_single_batch"] = {"ms_per_step": t1 * 1e3, "agent_steps_per_s_per_gpu": E * N / t1,
                                             "note": "one env batch, state stays in L2 between steps; not the headline"}
        del g1
    # persistent multi-step mode (flock_step_n): state stays in registers, no per-step HBM traffic
    if envs[0].tiled is False and rank == 0:
        T = 256
        envs[1].step_n(T, DT)
        torch.cuda.synchronize(device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        envs[1].step_n(T, DT)
        ev1.record()
        torch.cuda.synchronize(device)
        extra["step_n_persistent"] = {"steps_per_launch": T, "agent_steps_per_s_per_gpu": E * N * T / (ev0.elapsed_time(ev1) * 1e-3),
                                      "note": "in-kernel Philox actions, state in registers; FP32-issue bound, no per-step HBM traffic"}

    # the same kernel on ONE big batch (as many envs as the whole ring): shows what the kernel
    # sustains once a launch carries enough bytes to leave the launch-latency regime
    if rank == 0 and not args.no_sweep:
        from marl_range_flocking_b200 import VecEnv
        peak_, _ = _peaks()
        big_E = E * ring
        big = VecEnv(w["variant"], big_E, N, w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=1,
                     device=device, **w["kw"], **w.get("env_kw", {}))
        big.reset()
        big_act = [big.random_actions(i) for i in range(2)]
        for i in range(3):
            big.step(big_act[i & 1], DT)
        torch.cuda.synchronize(device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if not big.tiled else 3
        ev0.record()
        for i in range(reps):
            big.step(big_act[i & 1], DT)
        ev1.record()
        torch.cuda.synchronize(device)
        t_big = ev0.elapsed_time(ev1) * 1e-3 / reps
        gbs = big_E * N * w["bytes"] / t_big / 1e9
        extra["large_batch"] = {"envs": big_E, "ms_per_step": t_big * 1e3, "agent_steps_per_s": big_E * N / t_big,
                                "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak_,
                                "note": "one launch over envs = E x ring; working set > L2, same fused kernel"}
        del big, big_act

    # optional: closed-loop rollout with the batched per-agent actors of the shared-critic DDPG learner
    # (BASELINE configs[2] "MADDPG actor rollout"; SURVEY 8f-2: the policy, not the env, bounds it)
    # The fused tcgen05 leg always runs for the uw workload (cfg3 IS the actor-rollout config); the PyTorch
    # legs it is compared with only under --policy actor (they take seconds).
    if rank == 0 and w["variant"] != "uwd" and (args.policy == "actor" or w["variant"] == "uw"):
        from marl_range_flocking_b200.policies import BatchedActors
        env = envs[0]
        obs = env.observation
        in_dims = obs[0, 0].numel()
        extra["actor_rollout"] = {}
        for dtype, name in (((torch.float32, "fp32"), (torch.bfloat16, "bf16")) if args.policy == "actor" else ()):
            actors = BatchedActors(N, in_dims, 400, 300, 2, device=device, dtype=dtype)
            with torch.no_grad():
                for _ in range(5):
                    obs, *_ = env.step(actors(obs), DT)
                torch.cuda.synchronize(device)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 100
                ev0.record()
                for _ in range(reps):
                    obs, *_ = env.step(actors(obs), DT)
                ev1.record()
                torch.cuda.synchronize(device)
            t = ev0.elapsed_time(ev1) * 1e-3 / reps
            extra["actor_rollout"][name] = {
                "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                "policy": f"{N} per-agent MLPs {in_dims}-400-300-2 (LayerNorm, ReLU, tanh) as baddbmm over the agent dim"}
        # the same closed loop with the fused tcgen05 actor kernel (flock_actor_forward): one policy launch
        # + one env launch per step, CUDA-graph replay, actions written straight into the env's input
        actors = BatchedActors(N, in_dims, 400, 300, 2, device=device)
        actors.pack_fused()
        acts = torch.empty(E, N, 2, device=device)
        flops = 2.0 * E * N * (in_dims * 400 + 400 * 300 + 300 * 2)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                actors.forward_fused(env.observation, out=acts)
                env.step(acts, DT)
            side.synchronize()
            reps, inner = 20, 50
            for label, with_env in (("policy_only", False), ("closed_loop", True)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(inner):
                        actors.forward_fused(env.observation, out=acts)
                        if with_env:
                            env.step(acts, DT)
                g.replay()
                side.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(side)
                for _ in range(reps):
                    g.replay()
                ev1.record(side)
                side.synchronize()
                t = ev0.elapsed_time(ev1) * 1e-3 / (reps * inner)
                extra["actor_rollout"]["fused_tcgen05_" + label] = {
                    "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t, "policy_tflops": flops / t / 1e12,
                    "frac_of_bf16_peak": flops / t / 1e12 / _bf16_peak(),
                    "policy": "flock_actor_forward: tcgen05.mma bf16 / fp32 accumulate, LayerNorm + ReLU + tanh fused, "
                              "one launch for all envs and agents"}
        torch.cuda.current_stream(device).wait_stream(side)

    # the reference's default main.py loop: v2 env + recurrent MADDPG actors (learners/maddpg_official_rnn/net.py);
    # closed loop policy + env step with the fused kernels (fp32 GRU front end + tcgen05 MLP) under CUDA-graph replay,
    # next to the same networks as PyTorch baddbmm's
    if rank == 0 and args.policy == "actor" and w["variant"] == "v2":
        from marl_range_flocking_b200.policies import BatchedRnnActors
        env = envs[0]
        k_obs = env.observation.shape[-1]
        extra["rnn_actor_rollout"] = {}
        for dtype, name in ((torch.float32, "pytorch_fp32"), (torch.bfloat16, "pytorch_bf16")):
            net = BatchedRnnActors(N, k_obs, device=device, dtype=dtype)
            hidden = net.init_hidden(E)
            obs = env.observation
            with torch.no_grad():
                for _ in range(5):
                    act, hidden = net(obs, hidden)
                    obs, *_ = env.step(act, DT)
                torch.cuda.synchronize(device)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 100
                ev0.record()
                for _ in range(reps):
                    act, hidden = net(obs, hidden)
                    obs, *_ = env.step(act, DT)
                ev1.record()
                torch.cuda.synchronize(device)
            t = ev0.elapsed_time(ev1) * 1e-3 / reps
            extra["rnn_actor_rollout"][name] = {"ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t}
        net = BatchedRnnActors(N, k_obs, device=device)
        net.pack_fused()
        hidden = net.init_hidden(E)
        acts = torch.empty(E, N, 2, device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                net.forward_fused(env.observation, hidden, out=acts, hidden_out=hidden)
                env.step(acts, DT)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            reps, inner = 20, 50
            with torch.cuda.graph(g, stream=side):
                for _ in range(inner):
                    net.forward_fused(env.observation, hidden, out=acts, hidden_out=hidden)
                    env.step(acts, DT)
            g.replay()
            side.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(side)
            for _ in range(reps):
                g.replay()
            ev1.record(side)
            side.synchronize()
        torch.cuda.current_stream(device).wait_stream(side)
        t = ev0.elapsed_time(ev1) * 1e-3 / (reps * inner)
        extra["rnn_actor_rollout"]["fused_closed_loop"] = {
            "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
            "policy": f"{N} per-agent recurrent actors {k_obs}-32-GRU32-400-300-2: flock_rnn_actor_forward (fp32 GRU front kernel + "
                      "tcgen05 MLP kernel) + one env launch per step, CUDA-graph replay"}

    # VDN action selection next to the discrete env (BASELINE configs[3]): the per-agent Q networks of
    # learners/vdn/net.py (recurrent, as the reference trains them) + per-env epsilon-greedy, closed loop with the
    # env step; fused fp32 kernel (flock_qnet_forward) vs the same networks as PyTorch baddbmm's
    if rank == 0 and w["variant"] == "uwd":
        from marl_range_flocking_b200.policies import BatchedQNet
        env = envs[0]
        k_obs = env.observation.shape[-1]
        qn = BatchedQNet(N, k_obs, k_obs, recurrent=True, device=device)
        extra["vdn_rollout"] = {}
        legs = [("fused_fp32", True)] + ([("pytorch_fp32", False)] if args.policy == "actor" else [])
        for label, fused in legs:
            hidden = qn.init_hidden(E)
            obs = env.observation
            with torch.no_grad():
                def one(t, obs, hidden):
                    if fused:
                        act, hidden = qn.sample_action_fused(obs, hidden, 0.1, step=t, seed=11)
                    else:
                        act, hidden = qn.sample_action(obs, hidden, 0.1)
                    obs, *_ = env.step(act, DT)
                    return obs, hidden
                for t in range(5):
                    obs, hidden = one(t, obs, hidden)
                torch.cuda.synchronize(device)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 200
                ev0.record()
                for t in range(reps):
                    obs, hidden = one(5 + t, obs, hidden)
                ev1.record()
                torch.cuda.synchronize(device)
            t = ev0.elapsed_time(ev1) * 1e-3 / reps
            extra["vdn_rollout"][label] = {
                "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                "policy": f"{N} per-agent recurrent Q networks {k_obs}-64-32-GRU32-{k_obs} + per-env epsilon-greedy, "
                          + ("one fused fp32 launch (flock_qnet_forward)" if fused else "PyTorch baddbmm over the agent dim"),
                "launch": "eager (one policy launch + one env launch per step)"}

    if rank == 0:
        peak, peak_src = _peaks()
        per_launch_s = ms_max * 1e-3 / steps
        alg_bytes = E * N * w["bytes"]
        achieved = alg_bytes / per_launch_s / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": TRAFFIC_PER_LAUNCH.get(args.workload), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "achieved = algorithmic bytes / mean launch-to-launch time of the timed region"}
        if envs[0].tiled:
            # large swarms are FP32-pipe bound, not HBM bound: E*N*(N-1) pair evaluations x 9
            # FP32 operations (no FMA: parity forbids contraction) against 128 lanes x SMs x max clock
            # The v2 sensing kernel prunes exactly (box test against the warp's k-th-distance bound), so
            # the flops it needs are those of the pairs it evaluates: counted on a separate untimed env.
            props = torch.cuda.get_device_properties(device)
            all_pairs = E * N * (N - 1) * 9.0
            frac_eval = pairs_evaluated_fraction(w, E, device)
            flops = all_pairs * frac_eval
            peak_tf = props.multi_processor_count * 128 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12
            ach_tf = flops / per_launch_s / 1e12
            roof = {"bound": "fp32", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                    "traffic": TRAFFIC_PER_LAUNCH.get(args.workload),
                    "peak_source": "SMs x 128 FP32 lanes x max SM clock, one non-fused FP32 op per lane per clock",
                    "algorithmic_flops_per_launch": flops, "pairs_evaluated_frac": frac_eval,
                    "all_pairs_equivalent": {"flops_per_launch": all_pairs, "achieved": all_pairs / per_launch_s / 1e12,
                                             "frac": all_pairs / per_launch_s / 1e12 / peak_tf,
                                             "note": "what an all-pairs scan (the reference's cdist) would need for this step rate"},
                    "note": "time = whole step (integrate pre-pass + sensing kernel), flops = 9 per evaluated pair",
                    "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak}}
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "envs_per_gpu": E, "agents": N, "k": w["k"],
                       "l2_policy": f"inputs larger than L2: ring of {ring} independent env batches "
                                    f"({ring * bytes_per_batch / 2**20:.0f} MiB), step s touches batch s % {ring}",
                       "launch": ("CUDA graph replay, integrate pre-pass + sensing kernel per step (+ row-order refresh "
                                  "every 16 steps), single stream") if envs[0].tiled else
                                 "CUDA graph replay, one fused kernel per step, single stream"},
            "roofline": roof,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": e2e_api, "sync_per_step_value": e2e_sync_value,
                    "pipelined_value": e2e_pipe_value},
            "gpu_launches": launches * world,
            "stats_allreduce": {"backend": "nccl" if use_dist else "none", "episodes": int(stats[0].item())},
        }
        line.update(extra)
        if world == 1 and not args.no_cpu:
            r = cpu_run(w, 30, 3, budget_s=12.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "agent-steps/s", "cores": r["threads"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the committed
# ncu --set full captures (profiles/r01_cfg2_step_small_raw.txt, r01_cfg5_step_tiled_raw.txt); the
# result stores of one launch are still in L2 when the capture ends, so writes read ~0.
TRAFFIC_PER_LAUNCH = {"cfg2": 883200, "cfg5": 6857472}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20480)
    ap.add_argument("--warmup", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ring", type=int, default=0, help="number of env batches in the L2-defeating ring (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-batch roofline leg")
    ap.add_argument("--policy", default="none", choices=["none", "actor"],
                    help="extra leg: closed-loop rollout with batched per-agent actors (policies.BatchedActors)")
    ap.add_argument("--no-index", action="store_true",
                    help="do not track neighbour indices (values-only selection); default for uw / uwd, where the "
                         "reference discards them, opt-in for v2, where the reference keeps `nearest_neighbors`")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.no_index:
        w["env_kw"] = dict(w.get("env_kw", {}), track_neighbors=False)
        w["desc"] += " [neighbour indices not tracked]"
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_gpu(args, w)


if __name__ == "__main__":
    main()
