set -x
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r02/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest4.log
tail -15 gpurun_out/r02/pytest4.log
python - > gpurun_out/r02/reset_timing.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from marl_range_flocking_b200 import VecEnv
env = VecEnv("v2", 64, 2048, 8, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=1)
for label, mask in (("all 64 envs", None), ("1 env", torch.zeros(64, dtype=torch.bool, device="cuda").index_fill_(0, torch.tensor([5], device="cuda"), True))):
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
cat gpurun_out/r02/reset_timing.txt
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg3 --no-configs --no-cpu --no-sweep > gpurun_out/r02/bench4_cfg3.json 2> gpurun_out/r02/bench4_cfg3.err; echo rc=$?
