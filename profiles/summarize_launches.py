"""Aggregate an ncu --csv launch log (gpu__time_duration.sum per launch) into a per-kernel table."""
import csv
import collections
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    val = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    short = re.sub(r"<.*", "", name)[:70]
    agg[short][0] += 1
    agg[short][1] += val * scale
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
print(f"{'kernel':72s} {'n':>5s} {'us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:72s} {n:5d} {us:10.1f} {100*us/tot:6.1f}%")
