"""Developer check of the tensor-core GRU front ends (csrc/flock_gru_tc.cu): accuracy against the fp32 PyTorch
modules and timing against the fp32 CUDA-core kernels. usage: python tools/gru_tc_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedQNet, BatchedRnnActors

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                fn()
        g.replay()
        s.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(s)
        for _ in range(reps // 10):
            g.replay()
        ev1.record(s)
        s.synchronize()
    return ev0.elapsed_time(ev1) / (reps // 10 * 10) * 1e3


for E, N, n_obs, A in ((256, 2, 4, 4), (8192, 16, 4, 4)):
    torch.manual_seed(1)
    qn = BatchedQNet(N, n_obs, A, recurrent=True, device=dev)
    obs = torch.rand(E, N, n_obs, device=dev) * 7
    hid = torch.randn(E, N, 32, device=dev) * 0.5
    with torch.no_grad():
        q_ref, h_ref = qn(obs, hid)
    for impl in ("fp32", "tc"):
        q, h = qn.forward_fused(obs, hid, impl=impl)
        torch.cuda.synchronize()
        print(f"qnet E={E} N={N} {impl}: max|dq| {(q - q_ref).abs().max().item():.2e} max|dh| {(h - h_ref).abs().max().item():.2e}", flush=True)
    out = torch.empty(E, N, device=dev)
    for impl in ("fp32", "tc"):
        t = timeit(lambda: qn.sample_action_fused(obs, hid, 0.1, step=1, seed=2, out=out, hidden_out=hid, impl=impl))
        print(f"qnet E={E} N={N} {impl}: {t:.1f} us per launch", flush=True)

for E, N, n_obs in ((256, 2, 4), (4096, 10, 4), (4096, 32, 12)):
    torch.manual_seed(2)
    net = BatchedRnnActors(N, n_obs, device=dev)
    net.pack_fused()
    obs = torch.rand(E, N, n_obs, device=dev) * 14
    hid = torch.randn(E, N, 32, device=dev) * 0.5
    with torch.no_grad():
        a_ref, h_ref = net(obs, hid)
    for impl in ("fp32", "tc"):
        a, h = net.forward_fused(obs, hid, impl=impl)
        torch.cuda.synchronize()
        print(f"rnn E={E} N={N} {impl}: max|da| {(a - a_ref).abs().max().item():.2e} max|dh| {(h - h_ref).abs().max().item():.2e}", flush=True)
    acts = torch.empty(E, N, 2, device=dev)
    for impl in ("fp32", "tc"):
        t = timeit(lambda: net.forward_fused(obs, hid, out=acts, hidden_out=hid, impl=impl))
        print(f"rnn E={E} N={N} {impl}: {t:.1f} us per policy step", flush=True)
