set -x
mkdir -p gpurun_out/r02
(nproc; lscpu | head -30; free -g; nvidia-smi topo -m; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv) > gpurun_out/r02/host_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 -x -p no:cacheprovider > gpurun_out/r02/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest1.log
tail -5 gpurun_out/r02/pytest1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_short.json 2> gpurun_out/r02/bench_short.err; echo "rc=$?"
timeout 600 python bench.py --steps 20480 --warmup 1024 --no-configs --no-cpu --no-sweep > gpurun_out/r02/bench_long.json 2> gpurun_out/r02/bench_long.err; echo "rc=$?"
for g in 1 2; do FLOCK_FORCE_G=$g timeout 300 python bench.py --steps 1024 --warmup 64 --no-configs --no-cpu --no-sweep > gpurun_out/r02/bench_forceG$g.json 2> gpurun_out/r02/bench_forceG$g.err; done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02/bench_ref.json 2> gpurun_out/r02/bench_ref.err; echo "rc=$?"
