"""Phase timing of the tensor-core GRU front-end kernel (FLOCK_GRU_TIMING=1 makes the first launch print the mean
clock64 phase lengths). usage: python tools/gru_tc_phases.py qnet|rnn [E] [N]"""
import os
import sys

os.environ["FLOCK_GRU_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200.policies import BatchedQNet, BatchedRnnActors

which = sys.argv[1] if len(sys.argv) > 1 else "qnet"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda:0")
obs = torch.rand(E, N, 4, device=dev) * 7
hid = torch.randn(E, N, 32, device=dev) * 0.5
if which == "qnet":
    net = BatchedQNet(N, 4, 4, recurrent=True, device=dev)
    net.sample_action_fused(obs, hid, 0.1, step=1, seed=2)
else:
    net = BatchedRnnActors(N, 4, device=dev)
    net.pack_fused()
    net.forward_fused(obs, hid)
torch.cuda.synchronize()
