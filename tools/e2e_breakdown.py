"""Where does the end-to-end (host buffers) step time go? Run on a GPU box."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from marl_range_flocking_b200 import VecEnv

env = VecEnv("v2", 4096, 10, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=1)
env.reset()
acts = [env.random_actions(i).cpu().pin_memory() for i in range(4)]
for i in range(50):
    env.step_host(acts[i % 4])
n = 2000
t0 = time.perf_counter()
for i in range(n):
    env.step_host(acts[i % 4])
t_py = (time.perf_counter() - t0) / n * 1e6
hb = env._host
lib, h = env.lib, env._h
stream = torch.cuda.current_stream().cuda_stream
args = [(h, a.data_ptr(), ctypes.c_float(0.1), None, hb["obs"].data_ptr(), hb["reward"].data_ptr(),
         hb["agent_done"].data_ptr(), hb["env_done"].data_ptr(), stream) for a in acts]
t0 = time.perf_counter()
for i in range(n):
    lib.flock_step_host(*args[i % 4])
t_raw = (time.perf_counter() - t0) / n * 1e6
# device time of the zero-copy kernel alone (events around async launches through step_host is not possible: it syncs)
d_acts = [a.cuda() for a in acts]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
for i in range(n):
    env.step(d_acts[i % 4])
ev1.record()
torch.cuda.synchronize()
t_dev = ev0.elapsed_time(ev1) / n * 1e3
t0 = time.perf_counter()
for i in range(n):
    torch.cuda.current_stream().synchronize()
t_sync = (time.perf_counter() - t0) / n * 1e6
print(f"step_host via VecEnv: {t_py:.1f} us | raw ctypes flock_step_host: {t_raw:.1f} us | "
      f"async device step (python loop, L2 resident): {t_dev:.1f} us | empty stream sync: {t_sync:.1f} us")
for mode in ("0",):
    os.environ["FLOCK_ZEROCOPY"] = mode
