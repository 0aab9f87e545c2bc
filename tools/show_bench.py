"""Print the key numbers of a bench.py JSON line (developer convenience)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(path, "no json", e)
        continue
    print("==", path)
    print({k: d[k] for k in ("value", "ms_per_step", "repeats", "gpu_launches", "steps") if k in d})
    if "roofline" in d:
        r = d["roofline"]
        print("roofline", r["bound"], "frac %.4f" % r["frac"], "achieved %.1f" % r["achieved"], r["unit"], "traffic", r.get("traffic"))
    if "method" in d:
        print(d["method"]["spread_ms"])
    if "e2e" in d:
        e = d["e2e"]
        print("e2e %.3e" % e["value"], e.get("value_is"), "sync %.3e" % e.get("sync_per_step", {}).get("value", 0), "pipe %.3e" % e.get("pipelined", {}).get("value", 0))
    if "cpu_baseline" in d:
        print("cpu", json.dumps(d["cpu_baseline"])[:1200])
    for k in ("concurrent_batches", "l2_resident_single_batch", "step_n_persistent", "rollout_n_streamed", "large_batch", "stats_allreduce", "clocks", "actor_rollout",
              "vdn_rollout", "rnn_actor_rollout"):
        if k in d:
            print(k, json.dumps(d[k])[:700])
    for n, c in d.get("policies", {}).items():
        print(n, "us %.1f" % c["us_per_step"], "value %.3e" % c["agent_steps_per_s"])
    for n, c in d.get("configs", {}).items():
        print(n, "ms %.5f" % c["ms_per_step"], "value %.3e" % c["value"], "frac %.4f" % c["roofline"]["frac"], "repeats", c["repeats"])
