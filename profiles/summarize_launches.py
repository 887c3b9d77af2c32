"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    unit = "ns"
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        agg.setdefault(name, []).append(v)
    sc = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: per-kernel device time (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':72s} {'n':>5s} {'avg_us':>9s} {'min_us':>9s} {'max_us':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:72s} {len(v):5d} {sum(v) / len(v) * sc:9.2f} {min(v) * sc:9.2f} {max(v) * sc:9.2f} {sum(v) / tot:7.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
