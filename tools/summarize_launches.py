"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
bygrid = collections.defaultdict(list)
tot = 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    us = v / 1000.0 if row["Metric Unit"] in ("ns", "nsecond") else v
    agg[name][0] += 1
    agg[name][1] += us
    tot += us
    if name.startswith("conv_tc"):
        bygrid[row.get("Grid Size", "")].append(us)
print(f"launches {sum(n for n, _ in agg.values())}  total {tot:.1f} us (serialised, cold-cache ncu replay times)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={n:4d} total={t:9.1f} us  avg={t / n:7.1f} us  share={t / tot:.3f}")
for g, v in sorted(bygrid.items()):
    print(f"conv_tc grid {g}: n={len(v)} min {min(v):.1f} avg {sum(v) / len(v):.1f} max {max(v):.1f} us")
