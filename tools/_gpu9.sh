set -x
O=gpurun_out/r02c
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg3 --no-configs --no-cpu > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg4 --no-configs --no-cpu --policy actor > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg5 --no-configs --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg2 --no-configs --no-cpu --no-sweep --policy actor > $O/bench_cfg2_policy.json 2> $O/bench_cfg2_policy.err; echo "rc=$?"
python tools/gru_tc_check.py > $O/gru_tc_check.txt 2>&1
python tools/variant_tiled_timing.py > $O/variant_tiled_timing.txt 2>&1
python tools/gru_tc_phases.py qnet 8192 16 > $O/gru_tc_phases_qnet.txt 2>&1
python tools/gru_tc_phases.py rnn 4096 10 > $O/gru_tc_phases_rnn.txt 2>&1
tools/umma_rate > $O/umma_rate.txt 2>&1
# ncu: launch list of the default bench command, then --set full of the kernels that changed this round
python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > $O/bench_forlist.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/cfg2_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > $O/ncu_list.log 2>&1
python tools/prof_step.py --workload cfg2 --mode step --launches 12 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 6 -c 3 -f -o $O/cfg2_step python tools/prof_step.py --workload cfg2 --mode step --launches 12 > $O/ncu_cfg2_step.log 2>&1
for w in qnet rnn; do python tools/_prof_policy.py $w >> $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_gru_tc_kernel -s 3 -c 2 -f -o $O/gru_tc_$w python tools/_prof_policy.py $w > $O/ncu_gru_tc_$w.log 2>&1; done
for w in uw2048 uwd2048; do python tools/_prof_policy.py $w >> $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_pruned -s 20 -c 2 -f -o $O/pruned_$w python tools/_prof_policy.py $w > $O/ncu_pruned_$w.log 2>&1; done
ls -la $O
