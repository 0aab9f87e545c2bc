"""Fraction of the all-pairs work the pruned large-swarm kernel evaluates, and its speed: eager single
steps on one batch, then (like bench.py) CUDA-graph replay over a ring of batches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200 import VecEnv

E, N, k = 64, 2048, 8
ring = int(sys.argv[1]) if len(sys.argv) > 1 else 12


count = os.environ.get("FLOCK_COUNT", "1") != "0"


def make(seed):
    env = VecEnv("v2", E, N, k, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=seed)
    env.reset()
    if count:
        env.pairs_evaluated()      # the first query switches the counter on
    return env, [env.random_actions(i) for i in range(2)]


env, acts = make(3)
for t in range(40):
    env.pairs_evaluated(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    env.step(acts[t & 1])
    ev1.record()
    torch.cuda.synchronize()
    pe = env.pairs_evaluated()
    if t < 4 or t % 8 == 0:
        print(f"eager step {t:3d}: {ev0.elapsed_time(ev1) * 1e3:8.1f} us, evaluated {pe / (E * N * N):.3f} of all pairs")

envs = [make(100 + r) for r in range(ring)]
steps = 24 * ring
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(steps):
        e, a = envs[s % ring]
        e.step(a[(s // ring) & 1])
for rep in range(4):
    for e, _ in envs:
        e.pairs_evaluated(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    g.replay()
    ev1.record()
    torch.cuda.synchronize()
    pe = sum(e.pairs_evaluated() for e, _ in envs)
    print(f"graph replay {rep}: {ev0.elapsed_time(ev1) * 1e3 / steps:8.1f} us/step over a ring of {ring}, "
          f"evaluated {pe / (steps * E * N * N):.3f} of all pairs")
