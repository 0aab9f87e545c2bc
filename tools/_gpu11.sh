set -x
O=gpurun_out/r02h
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"
for w in cfg3 cfg4 cfg5; do timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-configs --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; echo "rc=$?"; done
python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > $O/bench_forlist.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/cfg2_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --no-configs > $O/ncu_list.log 2>&1
for w in cfg2 cfg3 cfg4; do python tools/prof_step.py --workload $w --mode step --launches 12 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 6 -c 3 -f -o $O/${w}_step python tools/prof_step.py --workload $w --mode step --launches 12 > $O/ncu_${w}_step.log 2>&1; done
python tools/prof_step.py --workload cfg2 --mode large --launches 6 >> $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_small -s 3 -c 2 -f -o $O/cfg2_large python tools/prof_step.py --workload cfg2 --mode large --launches 6 > $O/ncu_cfg2_large.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
