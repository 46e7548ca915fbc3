"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
order = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    key = (re.sub(r"\(.*", "", row["Kernel Name"]), row.get("Grid Size", ""))
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
    order.append((key[0][:40], v))
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:3d}  {k[0][:70]} grid={k[1]}")
if len(sys.argv) > 2:
    for name, v in order[: int(sys.argv[2])]:
        print(f"   {v:9.1f} us  {name}")
