set -x
O=gpurun_out/r02d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg5 --no-configs --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?"
python tools/variant_tiled_timing.py > $O/variant_tiled_timing.txt 2>&1
for w in uw2048 uwd2048; do python tools/_prof_policy.py $w > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flock_step_pruned -s 20 -c 2 -f -o $O/pruned_$w python tools/_prof_policy.py $w > $O/ncu_pruned_$w.log 2>&1; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
