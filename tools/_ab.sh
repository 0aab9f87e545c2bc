# usage: tools/_ab.sh "<tags>" "<workloads>"   (tag "base" = the regular library)
mkdir -p gpurun_out/ab
for w in $2; do for t in $1; do
  if [ "$t" = base ]; then unset FLOCK_LIBRARY_PATH; else export FLOCK_LIBRARY_PATH=$PWD/marl_range_flocking_b200/_ab/libflock_$t.so; fi
  timeout 300 python bench.py --workload $w --steps 2048 --warmup 64 --no-cpu --no-configs > gpurun_out/ab/${w}_$t.json 2> gpurun_out/ab/${w}_$t.err
  python - gpurun_out/ab/${w}_$t.json $w $t <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    lb = d.get("large_batch", {}); ro = d.get("rollout_n_streamed", {}); l2 = d.get("l2_resident_single_batch", {})
    print(sys.argv[2], sys.argv[3], "step_us %.3f" % (d["ms_per_step"] * 1e3), "l2_us %.3f" % (l2.get("ms_per_step", 0) * 1e3), "rollout_us %.3f" % (ro.get("ms_per_step", 0) * 1e3),
          "large %.3e frac %.3f" % (lb.get("agent_steps_per_s", 0), lb.get("frac_of_hbm_peak", 0)), flush=True)
except Exception as e:
    print(sys.argv[2], sys.argv[3], "failed", e)
PY
done; done
