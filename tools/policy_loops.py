"""The three closed policy + env loops of bench.py's `policies` key, alone (developer A/B). usage: python tools/policy_loops.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
out = bench.policy_closed_loops(dev, 0)
for k, v in out.items():
    print(f"{k}: {v['us_per_step']:.2f} us per step")
