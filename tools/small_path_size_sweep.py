"""Step time of the small path over a range of batch sizes (four L2-resident env batches, graph replay): developer A/B of
build variants (FLOCK_LIBRARY_PATH), e.g. trigger placement / pre-wait prefetch. usage: python tools/small_path_size_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_range_flocking_b200 import VecEnv
for variant, N, k in (("v2", 10, 4), ("uwd", 16, 4), ("uw", 32, 3)):
    for E in (64, 512, 2048, 4096, 16384, 65536):
        kw = dict(track_velocities=False)
        if variant != "v2":
            kw["track_neighbors"] = False
        envs = [VecEnv(variant, E, N, k, 2.5 if variant == "v2" else 0.5, range_start=(0, 50 if variant == "v2" else 200), sensor_range=14.0, seed=3 + r, **kw) for r in range(4)]
        for e in envs:
            e.reset()
        acts = [e.random_actions(0) for e in envs]
        for i in range(8):
            envs[i & 3].step(acts[i & 3])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(64):
                envs[i & 3].step(acts[i & 3])
        g.replay()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            g.replay()
        ev1.record()
        torch.cuda.synchronize()
        print(f"{variant} N={N} E={E:6d} {ev0.elapsed_time(ev1) * 1e3 / 640:8.2f} us/step", flush=True)
