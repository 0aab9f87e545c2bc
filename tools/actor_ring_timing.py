"""Fused actor kernel: observation window ([E][N][12]) vs observation ring ([E][4][N][3] + head) as input.
usage: python tools/actor_ring_timing.py [phases]   (phases: FLOCK_ACTOR_TIMING table of the first launch)"""
import os
import sys

if len(sys.argv) > 1:
    os.environ["FLOCK_ACTOR_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedActors
from marl_range_flocking_b200.vec_env import ObsRing

E, N = 4096, 32
dev = torch.device("cuda:0")
a = BatchedActors(N, 12, 400, 300, 2, device=dev)
a.pack_fused()
ring = ObsRing(torch.rand(E, 4, N, 3, device=dev) * 7, torch.full((E,), 2, dtype=torch.int32, device=dev))
win = ring.window().reshape(E, N, 12).contiguous()
out = torch.empty(E, N, 2, device=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
for name, obs in (("window", win), ("ring", ring)):
    if which not in ("both", name):
        continue
    for _ in range(3):
        a.forward_fused(obs, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                a.forward_fused(obs, out=out)
        g.replay()
        s.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(s)
        for _ in range(10):
            g.replay()
        ev1.record(s)
        s.synchronize()
    print(f"{name}: {ev0.elapsed_time(ev1) / 200 * 1e3:.2f} us per launch")
a2 = a.forward_fused(win).clone()
a3 = a.forward_fused(ring)
print("equal outputs:", bool(torch.equal(a2, a3)))
