set -x
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r02/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest5.log
tail -15 gpurun_out/r02/pytest5.log
python - > gpurun_out/r02/reset_timing2.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from marl_range_flocking_b200 import VecEnv
env = VecEnv("v2", 64, 2048, 8, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=1)
one = torch.zeros(64, dtype=torch.bool, device="cuda"); one[5] = True
none = torch.zeros(64, dtype=torch.bool, device="cuda")
for label, mask in (("all 64 envs", None), ("1 env", one), ("no env (auto-reset idle cost)", none)):
    env.reset(mask=mask); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10): env.reset(mask=mask)
    ev1.record(); torch.cuda.synchronize()
    print(label, "reset: %.1f us" % (ev0.elapsed_time(ev1) * 100))
a = env.random_actions()
for _ in range(20): env.step(a, 0.1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(50): env.step(a, 0.1)
ev1.record(); torch.cuda.synchronize()
print("step: %.1f us" % (ev0.elapsed_time(ev1) * 20))
PY
cat gpurun_out/r02/reset_timing2.txt
