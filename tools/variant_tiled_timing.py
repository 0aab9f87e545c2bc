"""Step time of the three variants on the tiled path (N = 2048, E = 64, k = 8 / 3 / 4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200 import VecEnv

for variant, k in (("v2", 8), ("uw", 3), ("uwd", 4), ("v2", 4)):
    env = VecEnv(variant, 64, 2048, k, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=3, reset_collision_distance=0.05)
    env.reset()
    acts = [env.random_actions(i) for i in range(2)]
    for i in range(3):
        env.step(acts[i & 1])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(20):
        env.step(acts[i & 1])
    ev1.record()
    torch.cuda.synchronize()
    print(f"{variant:4s} k={k} N=2048 E=64: {ev0.elapsed_time(ev1) * 1e3 / 20:8.1f} us/step")
