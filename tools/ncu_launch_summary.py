"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            if row["Metric Unit"] in ("us", "usecond"):
                v *= 1000.0
            elif row["Metric Unit"] in ("ms", "msecond"):
                v *= 1e6
            agg[row["Kernel Name"][:78]].append(v)
    total = sum(sum(v) for v in agg.values())
    print(f"{'kernel':80s} {'n':>5s} {'mean_ns':>10s} {'share':>7s}")
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n:80s} {len(v):5d} {sum(v) / len(v):10.0f} {100 * sum(v) / total:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
