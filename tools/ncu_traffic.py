"""Record the DRAM traffic of a kernel from an `ncu --set full` report into profiles/ncu_traffic.json.

    python tools/ncu_traffic.py gpurun_out/r02/cfg2_step.ncu-rep cfg2 [kernel-name-regex]

bench.py reads `roofline.traffic` for a workload from that file (dram__bytes_read.sum + dram__bytes_write.sum per
launch, averaged over the captured launches), together with the kernel name, the report name and the git SHA of the
tree the capture was taken from -- so a stale number is visible as such instead of living on as a constant.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main(report, workload, pattern=None):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    ii = hdr.index("smsp__inst_executed.sum") if "smsp__inst_executed.sum" in hdr else None
    sel = [r for r in data if pattern is None or re.search(pattern, r[ni])]
    if not sel:
        raise SystemExit("no launch matches " + str(pattern))
    f = lambda r, i: float(r[i].replace(",", "")) * unit_scale(units[i])
    rd = sum(f(r, ri) for r in sel) / len(sel)
    wr = sum(f(r, wi) for r in sel) / len(sel)
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    table[workload] = {"kernel": sel[0][ni][:120], "launches": len(sel), "dram_read_bytes_per_launch": rd,
                       "dram_write_bytes_per_launch": wr, "dram_bytes_per_launch": rd + wr,
                       "gpu_time_us_cold_serialised": sum(float(r[ti].replace(",", "")) for r in sel) / len(sel) * (1e-3 if units[ti] in ("ns", "nsecond") else 1.0),
                       "warp_instructions_per_launch": None if ii is None else sum(float(r[ii].replace(",", "")) for r in sel) / len(sel),
                       "capture": os.path.basename(report), "git_sha": sha,
                       "note": "ncu --set full --clock-control none; result stores of a launch may still sit in L2 when the capture ends"}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(table[workload], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
