"""Summarise an `ncu --page source --csv --print-source sass` dump: stall reasons and hot SASS lines."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[start[0]]
    end = start[1] - 1 if len(start) > 1 else len(rows)
    body = [r for r in rows[start[0] + 1:end] if len(r) >= len(h)]
    ci = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: 0 for s in stalls}
    samples = inst = 0
    for r in body:
        for s in stalls:
            try:
                tot[s] += int(r[ci[s]])
            except ValueError:
                pass
        samples += int(r[ci["# Samples"]] or 0)
        inst += int(r[ci["Instructions Executed"]] or 0)
    print(rows[start[0] - 1][1] if start[0] else "")
    print("sass lines", len(body), "samples", samples, "warp-instructions", inst)
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {s:28s} {v:7d} {100.0 * v / max(samples, 1):5.1f}%")
    print("hot lines (samples, warp-inst, sass):")
    for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:top]:
        print(r[ci["# Samples"]].rjust(6), r[ci["Instructions Executed"]].rjust(8), " ", r[ci["Source"]][:120])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
