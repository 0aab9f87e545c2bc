import sys, torch
sys.path.insert(0, "/root/repo")
from marl_range_flocking_b200 import VecEnv
env = VecEnv("v2", 50, 10, 4, 0.5, range_start=(0, 100), sensor_range=9.0, seed=77)
env.reset()
torch.cuda.synchronize()
print("reset ok")
env.step_n(3, 0.1)
torch.cuda.synchronize()
print("step_n ok")
