"""Short driver for ncu captures: a few launches of one kernel of the hot path on a BASELINE workload.

    python tools/prof_step.py --workload cfg2 --mode step --launches 12
    python tools/prof_step.py --workload cfg3 --mode rollout --launches 4

The env ring is larger than L2 (as in bench.py), so every launch meets cold data.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--mode", default="step", choices=["step", "rollout", "large"])
ap.add_argument("--launches", type=int, default=12)
ap.add_argument("--rollout-steps", type=int, default=64)
args = ap.parse_args()
w = dict(bench.WORKLOADS[args.workload])
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ring, _ = bench.ring_size(w)
if args.mode == "large":
    from marl_range_flocking_b200 import VecEnv
    env = VecEnv(w["variant"], w["E"] * ring, w["N"], w["k"], w["cd"], range_start=w["rs"], sensor_range=w["sr"], seed=1, device=dev,
                 **w["kw"], **w.get("env_kw", {}))
    env.reset()
    acts = [env.random_actions(i) for i in range(2)]
    for i in range(args.launches):
        env.step(acts[i & 1], bench.DT)
else:
    envs, acts = bench.build_ring(w, w["E"], ring, dev, env_offset=0)
    if args.mode == "step":
        for s in range(args.launches):
            envs[s % ring].step(acts[s % ring][(s // ring) & 1], bench.DT)
    else:
        T = args.rollout_steps
        trajs = [envs[i].alloc_trajectory(T) for i in range(min(ring, args.launches))]
        a_T = [torch.stack([acts[i][t & 1] for t in range(T)]) for i in range(len(trajs))]
        for s in range(args.launches):
            i = s % len(trajs)
            envs[i].rollout_n(a_T[i], trajs[i], bench.DT)
torch.cuda.synchronize()
print("ok", args.workload, args.mode, args.launches)
