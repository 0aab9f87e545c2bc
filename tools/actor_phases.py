"""Phase timing of the fused actor kernel (FLOCK_ACTOR_TIMING=1 makes the first launch print the mean
clock64 phase lengths per CTA) plus a plain timing loop. usage: python tools/actor_phases.py [E] [N]"""
import os
import sys

os.environ["FLOCK_ACTOR_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedActors

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
a = BatchedActors(N, 12, 400, 300, 2, device=dev)
obs = torch.rand(E, N, 12, device=dev) * 7
out = torch.empty(E, N, 2, device=dev)
a.pack_fused()
torch.cuda.synchronize()
for _ in range(3):
    a.forward_fused(obs, out=out)      # the first call prints the phase table
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(50):
    a.forward_fused(obs, out=out)
ev1.record()
torch.cuda.synchronize()
t = ev0.elapsed_time(ev1) / 50 * 1e3
flops = 2.0 * E * N * (12 * 400 + 400 * 300 + 300 * 2)
print(f"E={E} N={N}: {t:.1f} us per launch, {flops / t / 1e6:.1f} TFLOP/s")
