set -x
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r02/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest2.log
tail -30 gpurun_out/r02/pytest2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench2_cfg2.json 2> gpurun_out/r02/bench2_cfg2.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg3 --no-configs > gpurun_out/r02/bench2_cfg3.json 2> gpurun_out/r02/bench2_cfg3.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg4 --no-configs > gpurun_out/r02/bench2_cfg4.json 2> gpurun_out/r02/bench2_cfg4.err; echo "rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02/bench2_ref.json 2> gpurun_out/r02/bench2_ref.err; echo "rc=$?"
