#!/usr/bin/env python
"""Throughput benchmark of the flocking-env hot path (agent-steps/s), see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...   # CPU arm: the UNMODIFIED reference step on the host cores
                                           # (oracle/_ref, one process per core), else the C port

A "step" is ONE pass of `step()` over one batch of E envs x N agents (random actions already resident
in HBM). To keep the working set larger than the 126 MB L2, the bench cycles over a ring of R
independent env batches (step s touches batch s % R), all on one stream, replayed from CUDA graphs so
that Python launch overhead is not what is measured. Exactly K steps are timed per repetition; when K
steps are shorter than a few milliseconds the repetition is run many times at rotating ring offsets
(each graph exec replayed once un-timed first) and the MEDIAN is reported, so a `--steps 20` run
reproduces the sustained figure of a `--steps 20480` run. Under torchrun every rank owns its own ring
on its own GPU (envs shard with no data-path collective; the only collective is the NCCL all-reduce
of the episode statistics after the timed region) -> weak scaling, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs: density-scaled synthetic worlds, random actions
WORKLOADS = {
    # name: variant, E per GPU, N, k, collision, range_start, sensor, extra VecEnv kwargs, bytes/agent-step
    "cfg2": dict(variant="v2", E=4096, N=10, k=4, cd=2.5, rs=(0, 50), sr=14.0, kw={}, bytes=53,
                 env_kw=dict(track_velocities=False),   # `velocities` is an optional extra of SURVEY 8d (+8 B), not in the 53 B
                 desc="gym_flock_v2 batched 4096 envs x 10 agents, k=4, random actions (BASELINE configs[1])"),
    "cfg3": dict(variant="uw", E=4096, N=32, k=3, cd=0.5, rs=(0, 200), sr=7.0, kw={}, bytes=41,
                 # the reference discards the indices in uw (gym_flock_uw.py:141-144); ring layout: only the new range row is
                 # written per step (41 B), consumers read the history in place
                 env_kw=dict(track_neighbors=False, track_velocities=False, obs_layout="ring"),
                 desc="gym_flock_uw 4096 envs x 32 agents, k=3, random actions, ring-buffer observation history (BASELINE configs[2])"),
    "cfg4": dict(variant="uwd", E=8192, N=16, k=4, cd=0.5, rs=(0, 100), sr=7.0,
                 kw=dict(reset_collision_distance=1.0), bytes=49,
                 env_kw=dict(track_neighbors=False, track_velocities=False),   # ... and in uw_discrete (gym_flock_uw_discrete.py:189-192)
                 desc="gym_flock_uw_discrete 8192 envs x 16 agents, k=4, random action ids, Philox actuation noise (BASELINE configs[3])"),
    "cfg5": dict(variant="v2", E=64, N=2048, k=8, cd=0.05, rs=(0, 2000), sr=100.0, kw={}, bytes=69,
                 env_kw=dict(track_velocities=False),
                 desc="gym_flock_v2 large swarm 64 envs x 2048 agents, k=8, tiled path with exact box-pruned k-NN (BASELINE configs[4])"),
}
L2_BYTES = 126 * 1024 * 1024
DT = 0.1


def config_of(name, w):
    """The `config` object: identical in the GPU arm and the reference arm."""
    return {"workload": name + ": " + w["desc"], "envs_per_gpu": w["E"], "agents": w["N"], "k": w["k"]}


def _peaks_file():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _bf16_peak():
    """dense bf16 TFLOP/s: measured cuBLAS burst figure of MEASURED_PEAKS.json, else the 2250 nominal"""
    return float(_peaks_file().get("bf16_tflops", 2250.0))


def _peaks():
    p = _peaks_file()
    if "hbm_gbs" in p:
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py together with the git SHA
    of the capture); None when no capture of this workload has been committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
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
# CPU arms
#   port:      oracle/flock_oracle.c = a C restatement of the reference step, OpenMP over envs
#   reference: the UNMODIFIED reference modules (byte-compiled into oracle/_ref by oracle/build_ref.py,
#              loaded under oracle/ref_shim.py), one worker process per host core, CPU only
# ---------------------------------------------------------------------------------------------------
def cpu_run(w, steps, warmup, budget_s, nthreads=None):
    """Time `steps` oracle steps on a bounded sample of the workload's envs; returns a dict."""
    from oracle import flock_oracle as fo

    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle sets the
    # OpenMP team size explicitly, so that default does not apply)
    threads = nthreads or _host_threads() or fo.max_threads()
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


def ref_worker(spec_json):
    """Body of one reference worker process (CPU only): `m` instances of the reference's MultiAgentEnv, stepped
    one after the other with random actions, `env.reset()` whenever `done[1]` (the loop of main.py:31-51).
    Prints one JSON line {"seconds", "steps", "m", "resets"}."""
    spec = json.loads(spec_json)
    import torch

    from oracle import ref_shim

    torch.set_num_threads(int(spec["torch_threads"]))
    torch.manual_seed(int(spec["seed"]))
    mod = ref_shim.load_reference(spec["variant"])
    v, N, k = spec["variant"], spec["N"], spec["k"]
    kw = dict(agents=N, k=k, collision_distance=spec["cd"], range_start=tuple(spec["rs"]), sensor_range=spec["sr"])

    def sample_action():
        if v == "v2":
            return torch.rand(N, 2) * 3.0 - 1.5          # action_space Box(-1.5, 1.5), gym_flock_v2.py:58
        if v == "uw":
            return torch.rand(N, 2) * 2.0 - 1.0          # Box(-1, 1), gym_flock_uw.py:57
        return torch.randint(0, k, (N,)).float()          # Discrete(k) as float ids, vdn/net.py:55-57

    def make():
        e = mod.MultiAgentEnv(**kw)
        e.reset()
        return e

    def step_all(envs):
        n_reset = 0
        for e in envs:
            _, _, done, _ = e.step(sample_action(), dt=DT)
            if done[1]:
                e.reset()
                n_reset += 1
        return n_reset

    envs = [make()]
    t0 = time.perf_counter()
    for _ in range(3):
        step_all(envs)
    est = (time.perf_counter() - t0) / 3
    total = spec["steps"] + spec["warmup"]
    m = int(max(1, min(spec["m_max"], spec["budget_s"] / max(est * total, 1e-9))))
    envs += [make() for _ in range(m - 1)]
    for _ in range(spec["warmup"]):
        step_all(envs)
    # start together: wait until every worker has finished its set-up and warm-up
    if spec.get("sync_dir"):
        open(os.path.join(spec["sync_dir"], f"ready{spec['index']}"), "w").close()
        deadline = time.time() + 120.0
        while len([f for f in os.listdir(spec["sync_dir"]) if f.startswith("ready")]) < spec["workers"] and time.time() < deadline:
            time.sleep(0.002)
    resets = 0
    t0 = time.perf_counter()
    for _ in range(spec["steps"]):
        resets += step_all(envs)
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "steps": spec["steps"], "m": m, "resets": resets}), flush=True)


def ref_available():
    from oracle import ref_shim
    return ref_shim.reference_available() or ref_shim.compiled_reference_available()


def ref_run(w, steps, warmup, budget_s, workers=None, torch_threads=1, m_max=None):
    """Time the UNMODIFIED reference on a bounded sample of the workload: `workers` processes (default: one per
    host core), each stepping its own reference env instances. A bench "step" advances every instance once.
    Returns None when the reference is not available on this machine."""
    if not ref_available():
        return None
    P = workers or _host_threads()
    if m_max is None:
        m_max = max(1, w["E"] // P)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS=str(torch_threads), MKL_NUM_THREADS=str(torch_threads))
    with tempfile.TemporaryDirectory(prefix="flock_ref_sync_") as sync_dir:
        procs = []
        for i in range(P):
            spec = dict(variant=w["variant"], N=w["N"], k=w["k"], cd=w["cd"], rs=list(w["rs"]), sr=w["sr"], steps=steps,
                        warmup=warmup, budget_s=budget_s, m_max=m_max, torch_threads=torch_threads, seed=i, index=i,
                        workers=P, sync_dir=sync_dir if P > 1 else None)
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", json.dumps(spec)],
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=ROOT, text=True))
        outs = []
        for p in procs:
            out, err = p.communicate(timeout=600)
            lines = [l for l in out.splitlines() if l.startswith("{")]
            if p.returncode != 0 or not lines:
                sys.stderr.write("reference worker failed:\n" + err[-1500:] + "\n")
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                return None
            outs.append(json.loads(lines[-1]))
    envs = sum(o["m"] for o in outs)
    seconds = max(o["seconds"] for o in outs)
    # every worker times its own `steps` passes over its own instances; the job is done when the slowest is
    value = sum(o["m"] * w["N"] * steps for o in outs) / seconds
    return dict(value=value, seconds=seconds, envs=envs, workers=P, torch_threads=torch_threads,
                resets=sum(o["resets"] for o in outs),
                sample=f"{steps} steps of {envs} of the {w['E']} envs x {w['N']} agents: the unmodified reference "
                       f"MultiAgentEnv.step (+ reset on done[1]) under oracle/ref_shim.py, CPU, {P} worker processes x "
                       f"{torch_threads} torch thread(s), random actions")


def ref_cfg1():
    """BASELINE configs[0] (v2, one env, N=10, k=4, main.py defaults), 1000 steps,
    with one torch thread and with the default thread count."""
    w = dict(WORKLOADS["cfg2"], E=1)
    out = {}
    for label, threads in (("threads_1", 1), ("threads_default", _host_threads())):
        r = ref_run(w, 1000, 50, budget_s=1e9, workers=1, torch_threads=threads, m_max=1)
        if r is None:
            return None
        out[label] = {"env_steps_per_s": r["value"] / w["N"], "agent_steps_per_s": r["value"], "torch_threads": threads}
    out["config"] = "BASELINE configs[0]: gym_flock_v2, 1 env, 10 agents, k=4, range (0,50), collision 2.5, sensor 14, 1000 random-action steps"
    out["host"] = {"os_cpu_count": os.cpu_count(), "usable_threads": _host_threads(), "cpu_model": _cpu_model()}
    return out


def run_reference(args, name, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 2000), min(args.warmup, 50)
    port = cpu_run(w, steps, warmup, budget_s=15.0)
    r = ref_run(w, steps, warmup, budget_s=40.0)
    if r is not None:
        kind, value, seconds, cores, sample = "reference", r["value"], r["seconds"], r["workers"], r["sample"]
        note = ("the unmodified reference step (PyTorch, CPU) compiled into oracle/_ref, one process per host core; the C "
                "port of the same step (oracle/flock_oracle.c, OpenMP, all host threads) is in cpu_baseline.port")
    else:
        kind, value, seconds, cores, sample = "port", port["value"], port["seconds"], port["threads"], port["sample"]
        note = "oracle/_ref (the byte-compiled reference) is absent on this machine: this arm times the C port of its step"
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": seconds / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(name, w),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "port": {"value": port["value"], "cores": port["threads"], "sample": port["sample"]},
                         "host": {"os_cpu_count": os.cpu_count(), "usable_threads": _host_threads(), "cpu_model": _cpu_model()}},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": note,
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


def ring_size(w, ring_arg=0):
    bytes_per_batch = w["E"] * w["N"] * (w["bytes"] + 40)           # + nn_idx, episode counters, staging
    return (ring_arg or max(2, -(-int(1.25 * L2_BYTES) // bytes_per_batch))), bytes_per_batch


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


def timed_graph_steps(envs, acts, steps, warmup, device, barrier, reduce_max, target_steps=8192):
    """Device time per step of the sustained step stream (max over ranks).

    The unit of timing is one repetition = exactly `steps` steps. A short repetition cannot be timed alone: measured
    on B200 (profiles/README.md, round 2), a 20-step graph bracketed by its own event pair reads 3.55 us per step
    where the sustained stream runs at 3.02, because every isolated graph launch carries ~10 us of launch head. So
    `repeats` = ceil(target_steps / steps) repetitions are enqueued back to back as ONE contiguous stream (CUDA
    graphs of ~1000 steps continuing the ring where the previous one stopped, every graph exec replayed un-timed
    first) and timed as one region; this is done three times and the median region is reported. `--steps 20`
    therefore measures the same thing as `--steps 20480`. Returns ms per `steps` steps and the bookkeeping."""
    import torch

    R = len(envs)
    repeats = max(1, -(-target_steps // steps))
    total = steps * repeats
    chunk = max(R, (1024 // R) * R)               # a multiple of the ring: consecutive replays continue the ring
    full, rem = divmod(total, chunk)
    g_full, n_full = capture(envs, acts, chunk, 0) if full else (None, 0)
    g_rem, n_rem = capture(envs, acts, rem, (full * chunk) % R) if rem else (None, 0)

    def stream_once():
        for _ in range(full):
            g_full.replay()
        if g_rem is not None:
            g_rem.replay()

    for _ in range(max(1, -(-warmup // total))):   # warm-up: >= `warmup` steps, through the SAME graph execs
        stream_once()
    torch.cuda.synchronize(device)
    regions = []
    for _ in range(3):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        ev0.record()
        stream_once()
        ev1.record()
        torch.cuda.synchronize(device)
        barrier()
        regions.append(reduce_max(ev0.elapsed_time(ev1)))
    regions.sort()
    ms_total = regions[1]
    return dict(ms=ms_total / repeats, repeats=repeats, launches=(full * n_full + n_rem), launches_per_rep=(full * n_full + n_rem) // repeats,
                spread_ms={"regions_ms": regions, "per_step_us": [r / total * 1e3 for r in regions]},
                mode=f"{repeats} repetition(s) of exactly {steps} steps enqueued back to back as one contiguous stream of "
                     f"{total} steps ({full} replays of a {chunk}-step CUDA graph" + (f" + one {rem}-step graph" if rem else "")
                     + "; all graph execs replayed un-timed first), one event pair around the stream, median of 3 such regions, "
                       "max over ranks")


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
    d2h = env._obs_buf.numel() * 4 + E * N * 4 + E * N + E
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


def policy_closed_loops(device, rank):
    """Compact closed-loop legs for the default line: one fused policy launch + one env launch per step under CUDA-graph
    replay, one env batch (SURVEY 8f-2 / BASELINE configs[2], configs[3] and the reference's main.py loop)."""
    import torch
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.policies import BatchedActors, BatchedQNet, BatchedRnnActors

    def graph_time(step, inner=50, reps=10):
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                step()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(inner):
                    step()
            g.replay()
            side.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(side)
            for _ in range(reps):
                g.replay()
            ev1.record(side)
            side.synchronize()
        torch.cuda.current_stream(device).wait_stream(side)
        return ev0.elapsed_time(ev1) * 1e-3 / (reps * inner)

    def make(name, **over):
        w = WORKLOADS[name]
        kw = dict(w.get("env_kw", {}))
        kw.update(over)
        env = VecEnv(w["variant"], w["E"] if name != "cfg4" else 4096, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"],
                     seed=0xB0B, env_offset=rank * w["E"], device=device, **w["kw"], **kw)
        env.reset()
        return env

    out = {}
    # recurrent MADDPG actors of main.py on the v2 env (cfg2 shape)
    env = make("cfg2")
    E, N = env.num_envs, env.num_particles
    net = BatchedRnnActors(N, env.k, device=device)
    net.pack_fused()
    hidden, abuf = net.init_hidden(E), torch.empty(E, N, 2, device=device)

    def rnn_step():
        net.forward_fused(env.observation, hidden, out=abuf, hidden_out=hidden)
        env.step(abuf, DT)
    t = graph_time(rnn_step)
    out["rnn_actor_v2_4096x10"] = {"us_per_step": t * 1e6, "agent_steps_per_s": E * N / t,
                                   "policy": "recurrent MADDPG actors (fce + GRUCell on tcgen05 with split fp16 operands, MLP on tcgen05 bf16)"}
    del env, net
    # shared-critic DDPG actors on the uw env (cfg3 shape, window layout: the faster closed loop)
    env = make("cfg3", obs_layout="window")
    E, N = env.num_envs, env.num_particles
    actors = BatchedActors(N, env.obs_hist * env.k, 400, 300, 2, device=device)
    actors.pack_fused()
    abuf2 = torch.empty(E, N, 2, device=device)

    def actor_step():
        actors.forward_fused(env.observation, out=abuf2)
        env.step(abuf2, DT)
    t = graph_time(actor_step)
    out["ddpg_actor_uw_4096x32"] = {"us_per_step": t * 1e6, "agent_steps_per_s": E * N / t,
                                    "policy": "shared-critic DDPG actors 12-400-300-2 (tcgen05 bf16, LayerNorm / ReLU / tanh fused)"}
    del env, actors
    # VDN recurrent Q networks + epsilon-greedy on the discrete env (cfg4 variant at 4096 x 16)
    env = make("cfg4")
    E, N = env.num_envs, env.num_particles
    qn = BatchedQNet(N, env.k, env.k, recurrent=True, device=device)
    hidden, abuf3 = qn.init_hidden(E), torch.empty(E, N, device=device)

    def vdn_step():
        qn.sample_action_fused(env.observation, hidden, 0.1, step=0, seed=11, out=abuf3, hidden_out=hidden, counters=env.noise_counters)
        env.step(abuf3, DT)
    t = graph_time(vdn_step)
    out["vdn_qnet_uwd_4096x16"] = {"us_per_step": t * 1e6, "agent_steps_per_s": E * N / t,
                                   "policy": "recurrent VDN Q networks + per-env epsilon-greedy (tcgen05, split fp16 operands)"}
    return out


def hbm_roofline(w, E, N, per_launch_s, name):
    peak, peak_src = _peaks()
    alg_bytes = E * N * w["bytes"]
    achieved = alg_bytes / per_launch_s / 1e9
    tr = _traffic(name)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None if tr is None else tr.get("dram_bytes_per_launch"),
            "traffic_source": None if tr is None else {k: tr.get(k) for k in ("kernel", "capture", "git_sha")},
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "algorithmic_bytes_per_agent_step": w["bytes"],
            "note": "achieved = algorithmic bytes (SURVEY 8d) / median launch-to-launch time of the timed region"}


def fp32_roofline(w, E, N, per_launch_s, name, device, sm_max_mhz):
    """Large swarms are FP32-pipe bound, not HBM bound: E*N*(N-1) pair evaluations x 9 FP32 operations
    (no FMA: parity forbids contraction) against 128 lanes x SMs x max clock. The v2 sensing kernel prunes exactly
    (box test against the warp's k-th-distance bound), so the flops it needs are those of the pairs it evaluates:
    counted on a separate untimed env."""
    import torch

    props = torch.cuda.get_device_properties(device)
    all_pairs = E * N * (N - 1) * 9.0
    frac_eval = pairs_evaluated_fraction(w, E, device)
    flops = all_pairs * frac_eval
    peak_tf = props.multi_processor_count * 128 * (sm_max_mhz or 1965) * 1e6 / 1e12
    ach_tf = flops / per_launch_s / 1e12
    tr = _traffic(name)
    return {"bound": "fp32", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
            "traffic": None if tr is None else tr.get("dram_bytes_per_launch"),
            "peak_source": "SMs x 128 FP32 lanes x max SM clock, one non-fused FP32 op per lane per clock",
            "algorithmic_flops_per_launch": flops, "pairs_evaluated_frac": frac_eval,
            "all_pairs_equivalent": {"flops_per_launch": all_pairs, "achieved": all_pairs / per_launch_s / 1e12,
                                     "frac": all_pairs / per_launch_s / 1e12 / peak_tf,
                                     "note": "what an all-pairs scan (the reference's cdist) would need for this step rate"},
            "note": "time = whole step (integrate pre-pass + sensing kernel), flops = 9 per evaluated pair",
            "hbm": hbm_roofline(w, E, N, per_launch_s, name)}


def run_gpu(args, name, w):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    use_dist = world > 1
    # each rank's host loop (the e2e legs) on its own cores: 8 Python loops sharing one default affinity mask
    # migrate and collide (profiles/README.md, e2e scaling)
    pinned_cores = None
    if use_dist and os.environ.get("FLOCK_PIN_CORES", "1") != "0":
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local * per:(local + 1) * per]
            if mine:
                os.sched_setaffinity(0, mine)
                pinned_cores = [mine[0], mine[-1]]
        except (AttributeError, OSError):
            pass
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

        def reduce_max(v):
            t = torch.tensor([v], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
    else:
        def barrier():
            return None

        def reduce_max(v):
            return float(v)

    E, N = w["E"], w["N"]
    ring, bytes_per_batch = ring_size(w, args.ring)
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
    timing = timed_graph_steps(envs, acts, args.steps, args.warmup, device, barrier, reduce_max)
    clocks = sampler.stop()
    steps = args.steps
    ms_max = timing["ms"]
    value = world * E * N * steps / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C ABI calls (every rank, max time over ranks) ----
    e2e_steps = int(min(4096, max(args.steps, 512)))
    e2e_s, h2d, d2h = timed_e2e(envs[0], w, e2e_steps, 20, device)
    e2e_sync_value = world * E * N * e2e_steps / reduce_max(e2e_s)
    depth = min(len(envs), max(2, int(os.environ.get("FLOCK_E2E_DEPTH", "4"))))
    e2e_pipe_value = world * E * N * e2e_steps / reduce_max(timed_e2e_pipelined(envs[:depth], e2e_steps, 20, device))
    if e2e_pipe_value > e2e_sync_value:
        e2e_value, e2e_form = e2e_pipe_value, "pipelined"
    else:
        e2e_value, e2e_form = e2e_sync_value, "sync_per_step"
    e2e = {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "value_is": e2e_form,
           "sync_per_step": {"value": e2e_sync_value,
                             "api": "VecEnv.step_host -> flock_step_host: ONE env batch, pinned host buffers, stream "
                                    "synchronised every step (strictly the named config)"},
           "pipelined": {"value": e2e_pipe_value, "batches_in_flight": depth,
                         "api": f"VecEnv.step_host_async / wait_host -> flock_step_host_async: {depth} env batches of the "
                                "named size in flight, every step moves its actions from pinned host memory and its "
                                "results back to pinned host memory, wall clock"},
           "pinned_cores": pinned_cores}

    # ---- the one collective of the system: all-reduce of the episode statistics (NCCL over NVLink) ----
    # a 64-step auto-reset rollout per rank closes real episodes first, so the reduction moves non-zero counters
    from marl_range_flocking_b200 import VecEnv
    senv = VecEnv(w["variant"], min(E, 1024), N, w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=77,
                  env_offset=rank * E, device=device, auto_reset=True, **w["kw"], **w.get("env_kw", {}))
    senv.reset()
    for t in range(64):
        senv.step(senv.random_actions(), DT)
    stats = senv.stats_tensor().clone()
    local_episodes = int(stats[0].item())
    if use_dist:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats_line = {"backend": "nccl" if use_dist else "none", "episodes": int(stats[0].item()),
                  "episode_steps": int(stats[1].item()), "local_episodes_rank0": local_episodes,
                  "rollout": f"64 auto-reset steps of {min(E, 1024)} envs per rank before the reduction"}
    if use_dist and int(stats[0].item()) < local_episodes:
        raise SystemExit("stats all-reduce returned fewer episodes than this rank closed")
    del senv

    extra = {}
    peak_hbm, _ = _peaks()

    def event_time(fn, reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(reps):
            fn(i)
        ev1.record()
        torch.cuda.synchronize(device)
        return ev0.elapsed_time(ev1) * 1e-3 / reps

    # independent env batches in flight on several streams (informational: what a trainer with several replicas per GPU
    # gets -- the 4096-env launch fills 0.23 waves of the machine, so launches of different batches overlap; the headline
    # above is ONE stream). Same ring, same kernels, S graphs replayed concurrently, each over its own part of the ring.
    if rank == 0 and not envs[0].tiled and len(envs) >= 8:
        S = 4
        per = len(envs) // S
        streams = [torch.cuda.Stream(device=device) for _ in range(S)]
        graphs = []
        n_steps = max(per, (256 // per) * per)
        for si, st in enumerate(streams):
            sub_e, sub_a = envs[si * per:(si + 1) * per], acts[si * per:(si + 1) * per]
            st.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(st):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    for s_ in range(n_steps):
                        sub_e[s_ % per].step(sub_a[s_ % per][(s_ // per) & 1], DT)
                g.replay()
            graphs.append(g)
        torch.cuda.synchronize(device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 8
        ev0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream(device))
        for _ in range(reps):
            for st, g in zip(streams, graphs):
                with torch.cuda.stream(st):
                    g.replay()
        for st in streams:
            torch.cuda.current_stream(device).wait_stream(st)
        ev1.record()
        torch.cuda.synchronize(device)
        tc_ = ev0.elapsed_time(ev1) * 1e-3 / (reps * S * n_steps)
        gb = E * N * w["bytes"] / tc_ / 1e9
        extra["concurrent_batches"] = {"streams": S, "ms_per_step": tc_ * 1e3, "agent_steps_per_s_per_gpu": E * N / tc_,
                                       "achieved_GBps": gb, "frac_of_hbm_peak": gb / peak_hbm,
                                       "note": f"{S} independent env batches in flight ({S} streams, one CUDA graph each over its own "
                                               "quarter of the ring > L2); per-step time = region / all steps; not the headline"}
        del graphs

    # what a real rollout sees: ONE env batch stepped over and over, its state resident in L2 (informational: the
    # headline above cycles over a ring larger than L2, as the timing rules require)
    if rank == 0:
        g1, _ = capture(envs[:1], acts[:1], 512, 0)
        g1.replay()
        torch.cuda.synchronize(device)
        t1 = event_time(lambda i: g1.replay(), 4) / 512
        extra["l2_resident_single_batch"] = {"ms_per_step": t1 * 1e3, "agent_steps_per_s_per_gpu": E * N / t1,
                                             "note": "one env batch, state stays in L2 between steps; not the headline"}
        del g1
    # persistent multi-step mode (flock_step_n): state stays in registers, no per-step HBM traffic
    if envs[0].tiled is False and rank == 0:
        T = 256
        envs[1].step_n(T, DT)
        torch.cuda.synchronize(device)
        t = event_time(lambda i: envs[1].step_n(T, DT), 1)
        extra["step_n_persistent"] = {"steps_per_launch": T, "agent_steps_per_s_per_gpu": E * N * T / t,
                                      "note": "in-kernel Philox actions, state in registers; FP32-issue bound, no per-step HBM traffic"}
    # streamed rollout (flock_rollout_n): T steps per launch, actions read from and obs | reward | dones written to
    # time-major trajectory buffers every step -- the mode a replay writer uses
    if envs[0].tiled is False and rank == 0 and hasattr(envs[0], "rollout_n"):
        T = 128
        traj = envs[1].alloc_trajectory(T)
        acts_T = torch.stack([acts[1][t & 1] for t in range(T)])
        envs[1].rollout_n(acts_T, traj, DT)
        torch.cuda.synchronize(device)
        t = event_time(lambda i: envs[1].rollout_n(acts_T, traj, DT), 3)
        rb = envs[1].rollout_bytes_per_agent_step()
        extra["rollout_n_streamed"] = {
            "steps_per_launch": T, "ms_per_step": t / T * 1e3, "agent_steps_per_s_per_gpu": E * N * T / t,
            "algorithmic_bytes_per_agent_step": rb, "achieved_GBps": E * N * T * rb / t / 1e9,
            "frac_of_hbm_peak": E * N * T * rb / t / 1e9 / peak_hbm,
            "note": "flock_rollout_n: one launch = T steps; per step it reads that step's actions and writes that step's "
                    "obs | reward | agent_done | env_done into [T][E]... trajectory buffers; state stays in registers"}
        del traj, acts_T

    # the same kernel on ONE big batch (as many envs as the whole ring): shows what the kernel
    # sustains once a launch carries enough bytes to leave the launch-latency regime
    if rank == 0 and not args.no_sweep:
        big_E = E * ring
        big = VecEnv(w["variant"], big_E, N, w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=1,
                     device=device, **w["kw"], **w.get("env_kw", {}))
        big.reset()
        big_act = [big.random_actions(i) for i in range(2)]
        for i in range(3):
            big.step(big_act[i & 1], DT)
        torch.cuda.synchronize(device)
        t_big = event_time(lambda i: big.step(big_act[i & 1], DT), 20 if not big.tiled else 3)
        gbs = big_E * N * w["bytes"] / t_big / 1e9
        extra["large_batch"] = {"envs": big_E, "ms_per_step": t_big * 1e3, "agent_steps_per_s": big_E * N / t_big,
                                "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak_hbm,
                                "note": "one launch over envs = E x ring; working set > L2, same fused kernel"}
        del big, big_act

    # ---- the other BASELINE configs, compact (same method: ring > L2, graph replay, median of repetitions) ----
    if not args.no_configs:
        others = {}
        for oname in ("cfg3", "cfg4", "cfg5"):
            if oname == name:
                continue
            ow = dict(WORKLOADS[oname])
            oring, _ = ring_size(ow)
            oenvs, oacts = build_ring(ow, ow["E"], oring, device, env_offset=rank * ow["E"])
            ot = timed_graph_steps(oenvs, oacts, min(args.steps, 256), args.warmup, device, barrier, reduce_max,
                                   target_steps=2048 if not oenvs[0].tiled else 512)
            osteps = min(args.steps, 256)
            per = ot["ms"] * 1e-3 / osteps
            if rank == 0:
                if oenvs[0].tiled:
                    roof = fp32_roofline(ow, ow["E"], ow["N"], per, oname, device, clocks.get("sm_max_mhz"))
                    roof = {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "pairs_evaluated_frac")}
                else:
                    roof = hbm_roofline(ow, ow["E"], ow["N"], per, oname)
                    roof = {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "algorithmic_bytes_per_agent_step")}
                others[oname] = {"config": config_of(oname, ow), "ms_per_step": per * 1e3,
                                 "value": world * ow["E"] * ow["N"] / per, "unit": "agent-steps/s", "n_gpus": world,
                                 "steps": osteps, "repeats": ot["repeats"], "ring": oring, "roofline": roof}
            del oenvs, oacts
        if rank == 0:
            extra["configs"] = others
            extra["policies"] = policy_closed_loops(device, rank)

    # optional: closed-loop rollout with the batched per-agent actors of the shared-critic DDPG learner
    # (BASELINE configs[2] "MADDPG actor rollout"; SURVEY 8f-2: the policy, not the env, bounds it)
    # The fused tcgen05 leg always runs for the uw workload (cfg3 IS the actor-rollout config); the PyTorch
    # legs it is compared with only under --policy actor (they take seconds).
    if rank == 0 and w["variant"] != "uwd" and (args.policy == "actor" or w["variant"] == "uw"):
        from marl_range_flocking_b200.policies import BatchedActors
        env = envs[0]
        in_dims = env.obs_hist * env.k
        extra["actor_rollout"] = {}
        for dtype, pname in (((torch.float32, "fp32"), (torch.bfloat16, "bf16")) if args.policy == "actor" else ()):
            actors = BatchedActors(N, in_dims, 400, 300, 2, device=device, dtype=dtype)
            with torch.no_grad():
                for _ in range(5):
                    obs, *_ = env.step(actors(env.observation), DT)
                torch.cuda.synchronize(device)
                t = event_time(lambda i: env.step(actors(env.observation), DT), 100)
            extra["actor_rollout"][pname] = {
                "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                "policy": f"{N} per-agent MLPs {in_dims}-400-300-2 (LayerNorm, ReLU, tanh) as baddbmm over the agent dim"}
        # the same closed loop with the fused tcgen05 actor kernel (flock_actor_forward): one policy launch
        # + one env launch per step, CUDA-graph replay, actions written straight into the env's input.
        # Two env layouts: "ring" (the env writes 12 B per agent-step, the actor gathers its 4 history rows from 4
        # slots: slower policy, faster env) and "window" (the env shifts the 48-byte window, the actor reads it with
        # three 128-bit loads). Both are reported; the faster closed loop is `fused_tcgen05_closed_loop`.
        actors = BatchedActors(N, in_dims, 400, 300, 2, device=device)
        actors.pack_fused()
        abuf = torch.empty(E, N, 2, device=device)
        flops = 2.0 * E * N * (in_dims * 400 + 400 * 300 + 300 * 2)
        layouts = [("ring" if env.obs_ring else "window", env)]
        if env.obs_ring:
            wkw = dict(w.get("env_kw", {}), obs_layout="window")
            wenv = VecEnv(w["variant"], E, N, w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=0x5EED,
                          device=device, **w["kw"], **wkw)
            wenv.reset()
            layouts.append(("window", wenv))
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for lname, lenv in layouts:
                lobs = (lambda e=lenv: e.obs_handle) if lenv.obs_ring else (lambda e=lenv: e.observation)
                for _ in range(3):
                    actors.forward_fused(lobs(), out=abuf)
                    lenv.step(abuf, DT)
                side.synchronize()
                reps, inner = 20, 50
                for label, with_env in (("policy_only", False), ("closed_loop", True)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        for _ in range(inner):
                            actors.forward_fused(lobs(), out=abuf)
                            if with_env:
                                lenv.step(abuf, DT)
                    g.replay()
                    side.synchronize()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record(side)
                    for _ in range(reps):
                        g.replay()
                    ev1.record(side)
                    side.synchronize()
                    t = ev0.elapsed_time(ev1) * 1e-3 / (reps * inner)
                    extra["actor_rollout"][f"fused_tcgen05_{label}_{lname}"] = {
                        "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t, "policy_tflops": flops / t / 1e12,
                        "frac_of_bf16_peak": flops / t / 1e12 / _bf16_peak(), "obs_layout": lname,
                        "policy": "flock_actor_forward: tcgen05.mma bf16 / fp32 accumulate, LayerNorm + ReLU + tanh fused, "
                                  "one launch for all envs and agents"}
        torch.cuda.current_stream(device).wait_stream(side)
        for label in ("policy_only", "closed_loop"):
            best = min((v for k_, v in extra["actor_rollout"].items() if k_.startswith(f"fused_tcgen05_{label}_")),
                       key=lambda v: v["ms_per_step"])
            extra["actor_rollout"][f"fused_tcgen05_{label}"] = best

    # the reference's default main.py loop: v2 env + recurrent MADDPG actors (learners/maddpg_official_rnn/net.py);
    # closed loop policy + env step with the fused kernels (GRU front end + tcgen05 MLP) under CUDA-graph replay,
    # next to the same networks as PyTorch baddbmm's
    if rank == 0 and args.policy == "actor" and w["variant"] == "v2":
        from marl_range_flocking_b200.policies import BatchedRnnActors
        env = envs[0]
        k_obs = env.observation.shape[-1]
        extra["rnn_actor_rollout"] = {}
        for dtype, pname in ((torch.float32, "pytorch_fp32"), (torch.bfloat16, "pytorch_bf16")):
            net = BatchedRnnActors(N, k_obs, device=device, dtype=dtype)
            state = {"hidden": net.init_hidden(E)}

            def one(i, net=net, state=state):
                act, state["hidden"] = net(env.observation, state["hidden"])
                env.step(act, DT)
            with torch.no_grad():
                for i in range(5):
                    one(i)
                torch.cuda.synchronize(device)
                t = event_time(one, 100)
            extra["rnn_actor_rollout"][pname] = {"ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t}
        net = BatchedRnnActors(N, k_obs, device=device)
        net.pack_fused()
        hidden = net.init_hidden(E)
        abuf = torch.empty(E, N, 2, device=device)
        for impl, key, what in (("tc", "fused_closed_loop", "GRU front end on the tensor cores (tcgen05, split fp16 operands)"),
                                ("fp32", "fused_closed_loop_fp32_front", "fp32 CUDA-core GRU front kernel")):
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    net.forward_fused(env.observation, hidden, out=abuf, hidden_out=hidden, impl=impl)
                    env.step(abuf, DT)
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                reps, inner = 20, 50
                with torch.cuda.graph(g, stream=side):
                    for _ in range(inner):
                        net.forward_fused(env.observation, hidden, out=abuf, hidden_out=hidden, impl=impl)
                        env.step(abuf, DT)
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
            extra["rnn_actor_rollout"][key] = {
                "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                "policy": f"{N} per-agent recurrent actors {k_obs}-32-GRU32-400-300-2: {what} + tcgen05 MLP kernel + one env "
                          "launch per step, CUDA-graph replay"}

    # VDN action selection next to the discrete env (BASELINE configs[3]): the per-agent Q networks of
    # learners/vdn/net.py (recurrent, as the reference trains them) + per-env epsilon-greedy, closed loop with the
    # env step; fused kernels (flock_qnet_forward_tc / flock_qnet_forward) under CUDA-graph replay (the exploration draws take their step
    # from a device counter, so replays do not repeat them) vs the same networks as PyTorch baddbmm's (eager)
    if rank == 0 and w["variant"] == "uwd":
        from marl_range_flocking_b200.policies import BatchedQNet
        env = envs[0]
        k_obs = env.observation.shape[-1]
        qn = BatchedQNet(N, k_obs, k_obs, recurrent=True, device=device)
        extra["vdn_rollout"] = {}
        hidden = qn.init_hidden(E)
        abuf = torch.empty(E, N, device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        for impl, key, what in (("tc", "fused_tc", "one fused tensor-core launch (flock_qnet_forward_tc: tcgen05, split fp16 operands, fp32-level accuracy)"),
                                ("fp32", "fused_fp32", "one fused fp32 CUDA-core launch (flock_qnet_forward)")):
            with torch.cuda.stream(side), torch.no_grad():
                def fused_step():
                    qn.sample_action_fused(env.observation, hidden, 0.1, step=0, seed=11, out=abuf, hidden_out=hidden,
                                           counters=env.noise_counters, impl=impl)
                    env.step(abuf, DT)
                for _ in range(3):
                    fused_step()
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                reps, inner = 10, 50
                with torch.cuda.graph(g, stream=side):
                    for _ in range(inner):
                        fused_step()
                g.replay()
                side.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(side)
                for _ in range(reps):
                    g.replay()
                ev1.record(side)
                side.synchronize()
            t = ev0.elapsed_time(ev1) * 1e-3 / (reps * inner)
            extra["vdn_rollout"][key] = {
                "ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                "policy": f"{N} per-agent recurrent Q networks {k_obs}-64-32-GRU32-{k_obs} + per-env epsilon-greedy, {what}",
                "launch": "CUDA graph replay (policy launch + env launch per step)"}
        torch.cuda.current_stream(device).wait_stream(side)
        if args.policy == "actor":
            state = {"hidden": qn.init_hidden(E)}

            def one(i):
                act, state["hidden"] = qn.sample_action(env.observation, state["hidden"], 0.1)
                env.step(act, DT)
            with torch.no_grad():
                for i in range(5):
                    one(i)
                torch.cuda.synchronize(device)
                t = event_time(one, 200)
            extra["vdn_rollout"]["pytorch_fp32"] = {"ms_per_step": t * 1e3, "agent_steps_per_s": E * N / t,
                                                    "policy": "PyTorch baddbmm over the agent dim", "launch": "eager"}

    if rank == 0:
        per_launch_s = ms_max * 1e-3 / steps
        if envs[0].tiled:
            roof = fp32_roofline(w, E, N, per_launch_s, name, device, clocks.get("sm_max_mhz"))
        else:
            roof = hbm_roofline(w, E, N, per_launch_s, name)
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(name, w),
            "repeats": timing["repeats"],
            "method": {"timing": timing["mode"], "spread_ms": timing["spread_ms"],
                       "l2_policy": f"inputs larger than L2: ring of {ring} independent env batches "
                                    f"({ring * bytes_per_batch / 2**20:.0f} MiB), step s touches batch s % {ring}",
                       "launch": ("CUDA graph replay, integrate pre-pass + sensing kernel per step (+ row-order refresh "
                                  "every 16 steps), single stream") if envs[0].tiled else
                                 "CUDA graph replay, one fused kernel per step, single stream"},
            "roofline": roof,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": timing["launches"] * world,
            "gpu_launches_per_repetition": timing["launches_per_rep"] * world,
            "stats_allreduce": stats_line,
        }
        line.update(extra)
        if world == 1 and not args.no_cpu:
            port = cpu_run(w, 30, 3, budget_s=10.0)
            r = ref_run(w, 10, 2, budget_s=12.0)
            if r is not None:
                line["cpu_baseline"] = {"value": r["value"], "unit": "agent-steps/s", "cores": r["workers"], "kind": "reference",
                                        "sample": r["sample"]}
            else:
                line["cpu_baseline"] = {"value": port["value"], "unit": "agent-steps/s", "cores": port["threads"],
                                        "kind": "port", "sample": port["sample"]}
            line["cpu_baseline"]["port"] = {"value": port["value"], "cores": port["threads"], "sample": port["sample"]}
            line["cpu_baseline"]["reference_pytorch"] = ref_cfg1()
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--ref-worker":
        ref_worker(sys.argv[2])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ring", type=int, default=0, help="number of env batches in the L2-defeating ring (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-batch roofline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact cfg3 / cfg4 / cfg5 sub-keys")
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
        run_reference(args, args.workload, w)
    else:
        run_gpu(args, args.workload, w)


if __name__ == "__main__":
    main()
