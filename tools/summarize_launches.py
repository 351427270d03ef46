"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log by kernel name (+grid)."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((r["Kernel Name"].split("(")[0][:60], r.get("Grid Size", ""), ns))
tot = sum(r[2] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, grid, ns in rows:
    agg[name][0] += 1
    agg[name][1] += ns
print("launches %d  total %.3f ms" % (len(rows), tot / 1e6))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d  %9.3f ms  %5.1f%%" % (name, n, ns / 1e6, 100 * ns / tot))
if "--top" in sys.argv:
    print("--- slowest launches")
    for name, grid, ns in sorted(rows, key=lambda r: -r[2])[:40]:
        print("%-62s grid %-14s %9.3f ms" % (name, grid, ns / 1e6))
