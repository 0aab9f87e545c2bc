"""Dynamic warp-instruction count and stall samples per CUDA source line of one kernel.

Joins `ncu -i rep --page source --csv --print-source sass` (per-SASS-instruction counters, address
order) with `nvdisasm -g -c <cubin>` (the same instructions, with `//## File ... line N` markers) by
instruction offset.  usage: ncu_line_profile.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-substr> [top]
"""
import collections
import csv
import re
import sys


def static_lines(path, kernel_substr):
    """offset -> (file, line) for the kernel's SASS instructions. Inlined code is attributed to the
    innermost line marker, like ncu's source page."""
    out = {}
    in_k = False
    cur = ("?", 0)
    for ln in open(path):
        if ln.startswith(".text."):
            in_k = kernel_substr in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main(ncu_csv, sass, kernel_substr, top=40):
    rows = list(csv.reader(open(ncu_csv)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[start[0]]
    end = start[1] - 1 if len(start) > 1 else len(rows)
    body = [r for r in rows[start[0] + 1:end] if len(r) >= len(h)]
    ci = {n: i for i, n in enumerate(h)}
    base = int(body[0][ci["Address"]], 16)
    where = static_lines(sass, kernel_substr)
    inst = collections.Counter()
    samp = collections.Counter()
    for r in body:
        off = int(r[ci["Address"]], 16) - base
        key = where.get(off, ("?", 0))
        inst[key] += int(r[ci["Instructions Executed"]] or 0)
        samp[key] += int(r[ci["# Samples"]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"warp-instructions {ti}, samples {ts}")
    cache = {}
    for key, n in inst.most_common(top):
        f, l = key
        text = ""
        try:
            if f not in cache:
                cache[f] = open("marl_range_flocking_b200/csrc/" + f).read().split("\n")
            text = cache[f][l - 1].strip()[:88]
        except Exception:
            pass
        print(f"{100.0 * n / ti:5.1f}% inst {100.0 * samp[key] / max(ts, 1):5.1f}% samp  {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
