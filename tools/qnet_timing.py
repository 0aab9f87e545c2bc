"""Kernel time of the fused VDN action-selection kernel (CUDA-graph replay, no host overhead) next to the
eager call and the PyTorch module. usage: python tools/qnet_timing.py [E] [N]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedQNet

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
for rec in (True, False):
    net = BatchedQNet(N, 4, 4, recurrent=rec, device=dev)
    obs = torch.rand(E, N, 4, device=dev) * 7
    hidden = net.init_hidden(E)
    for _ in range(3):
        net.sample_action_fused(obs, hidden, 0.1, step=1)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            h = hidden
            for t in range(20):
                act, h = net.sample_action_fused(obs, h, 0.1, step=t)
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
    t0 = time.perf_counter()
    for t in range(200):
        act, h = net.sample_action_fused(obs, hidden, 0.1, step=t)
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t0) / 200 * 1e6
    with torch.no_grad():
        for _ in range(3):
            net.sample_action(obs, hidden, 0.1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(200):
            net.sample_action(obs, hidden, 0.1)
        torch.cuda.synchronize()
    t_torch = (time.perf_counter() - t0) / 200 * 1e6
    flops = 2.0 * E * N * (4 * 64 + 64 * 32 + (6144 if rec else 0) + 32 * 4)
    print(f"recurrent={rec} E={E} N={N}: fused kernel {t_graph:.1f} us ({flops / t_graph / 1e6:.1f} TFLOP/s fp32), "
          f"fused eager call {t_eager:.1f} us, PyTorch sample_action {t_torch:.1f} us")
