"""Short driver for ncu captures of the policy / large-swarm kernels.
    python tools/_prof_policy.py qnet|rnn|uw2048|uwd2048"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200 import VecEnv
from marl_range_flocking_b200.policies import BatchedQNet, BatchedRnnActors

which = sys.argv[1]
dev = torch.device("cuda:0")
if which == "qnet":
    E, N = 8192, 16
    net = BatchedQNet(N, 4, 4, recurrent=True, device=dev)
    obs, hid = torch.rand(E, N, 4, device=dev) * 7, torch.randn(E, N, 32, device=dev) * 0.5
    out = torch.empty(E, N, device=dev)
    for i in range(6):
        net.sample_action_fused(obs, hid, 0.1, step=i, seed=2, out=out, hidden_out=hid)
elif which == "rnn":
    E, N = 4096, 10
    net = BatchedRnnActors(N, 4, device=dev)
    net.pack_fused()
    obs, hid = torch.rand(E, N, 4, device=dev) * 14, torch.randn(E, N, 32, device=dev) * 0.5
    out = torch.empty(E, N, 2, device=dev)
    for i in range(6):
        net.forward_fused(obs, hid, out=out, hidden_out=hid)
else:
    variant, k = ("uw", 3) if which == "uw2048" else ("uwd", 4)
    env = VecEnv(variant, 64, 2048, k, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=3, reset_collision_distance=0.05)
    env.reset()
    acts = [env.random_actions(i) for i in range(2)]
    for i in range(24):
        env.step(acts[i & 1])
torch.cuda.synchronize()
print("ok", which)
