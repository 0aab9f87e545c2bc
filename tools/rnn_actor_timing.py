"""Time of the fused recurrent MADDPG actor (CUDA-graph replay of front kernel + MLP kernel) next to the PyTorch module.
usage: python tools/rnn_actor_timing.py [E] [N] [n_obs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedRnnActors

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n_obs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
a = BatchedRnnActors(N, n_obs, device=dev)
obs = torch.rand(E, N, n_obs, device=dev) * 14
hidden = a.init_hidden(E)
acts = torch.empty(E, N, 2, device=dev)
for _ in range(3):
    a.forward_fused(obs, hidden, out=acts, hidden_out=hidden)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for t in range(20):
            a.forward_fused(obs, hidden, out=acts, hidden_out=hidden)
    g.replay()
    side.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(side)
    for _ in range(20):
        g.replay()
    ev1.record(side)
    side.synchronize()
t_graph = ev0.elapsed_time(ev1) / 400 * 1e3
torch.cuda.synchronize()
for dtype, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
    b = BatchedRnnActors(N, n_obs, device=dev, dtype=dtype)
    with torch.no_grad():
        for _ in range(3):
            b(obs, hidden)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(50):
            b(obs, hidden)
        ev1.record()
        torch.cuda.synchronize()
    print(f"PyTorch BatchedRnnActors {name}: {ev0.elapsed_time(ev1) / 50 * 1e3:.1f} us")
flops = 2.0 * E * N * (n_obs * 32 + 2 * 32 * 96 + 32 * 400 + 400 * 300 + 600)
print(f"E={E} N={N} n_obs={n_obs}: fused {t_graph:.1f} us per policy step ({flops / t_graph / 1e6:.1f} TFLOP/s)")
