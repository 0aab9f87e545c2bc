import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_range_flocking_b200 import VecEnv
for E in (8, 16, 24, 32, 48):
    env = VecEnv("v2", E, 2048, 8, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=3, tiled_mode=1)
    env.reset()
    acts = [env.random_actions(i) for i in range(2)]
    for i in range(5):
        env.step(acts[i & 1])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(32):
            env.step(acts[i & 1])
    g.replay()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5):
        g.replay()
    ev1.record()
    torch.cuda.synchronize()
    print(f"E={E:3d} {ev0.elapsed_time(ev1) * 1e3 / 160:8.1f} us/step", flush=True)
