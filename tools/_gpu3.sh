set -x
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r02/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest3.log
tail -8 gpurun_out/r02/pytest3.log
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg3 --no-configs --no-cpu > gpurun_out/r02/bench3_cfg3.json 2> gpurun_out/r02/bench3_cfg3.err; echo "rc=$?"
python tools/host_copy_bw.py > gpurun_out/r02/host_copy_bw_1gpu.json 2> gpurun_out/r02/host_copy_bw_1gpu.err
# ncu: launch list of the default bench command, then --set full of the three kernels
python tools/prof_step.py --workload cfg2 --mode step --launches 12 > gpurun_out/r02/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 6 -c 3 -o gpurun_out/r02/cfg2_step python tools/prof_step.py --workload cfg2 --mode step --launches 12 > gpurun_out/r02/ncu_cfg2_step.log 2>&1
python tools/prof_step.py --workload cfg3 --mode step --launches 12 >> gpurun_out/r02/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 6 -c 3 -o gpurun_out/r02/cfg3_step python tools/prof_step.py --workload cfg3 --mode step --launches 12 > gpurun_out/r02/ncu_cfg3_step.log 2>&1
python tools/prof_step.py --workload cfg2 --mode rollout --launches 4 >> gpurun_out/r02/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 2 -c 2 -o gpurun_out/r02/cfg2_rollout python tools/prof_step.py --workload cfg2 --mode rollout --launches 4 > gpurun_out/r02/ncu_cfg2_rollout.log 2>&1
python tools/prof_step.py --workload cfg2 --mode large --launches 6 >> gpurun_out/r02/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 3 -c 2 -o gpurun_out/r02/cfg2_large python tools/prof_step.py --workload cfg2 --mode large --launches 6 > gpurun_out/r02/ncu_cfg2_large.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > gpurun_out/r02/bench_forlist.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/cfg2_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > gpurun_out/r02/ncu_list.log 2>&1
ls -la gpurun_out/r02/
