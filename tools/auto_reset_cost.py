"""Per-step cost of auto-reset (fused into the step kernel for N <= 32) vs plain step, cfg2 shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200 import VecEnv

for auto in (False, True):
    envs = [VecEnv("v2", 4096, 10, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=r, auto_reset=auto) for r in range(44)]
    acts = []
    for e in envs:
        e.reset()
        acts.append([e.random_actions(i) for i in range(2)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s in range(1024):
            envs[s % 44].step(acts[s % 44][(s // 44) & 1], 0.1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20):
        g.replay()
    ev1.record()
    torch.cuda.synchronize()
    done_frac = float(torch.stack([e.dones[1].float().mean() for e in envs]).mean())
    print(f"auto_reset={auto}: {ev0.elapsed_time(ev1) * 1e3 / (20 * 1024):.2f} us/step, env_done fraction per step {done_frac:.3f}, "
          f"episodes closed {sum(e.stats()['episodes'] for e in envs)}")
    del envs, acts, g
