"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-name totals for the
LAST quarter of the launches (= the one timed step of `bench.py --steps 1 --warmup 3`) and the per-launch list
of the conv kernels in that step (grid, duration)."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((int(r["ID"]), r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""), v * scale))
n = len(rows)
print(f"total launches {n}")
# one step = launches after the warm-up steps; model construction adds a handful of launches up front
step = rows[-(n // 4):] if n >= 8 else rows
tot = defaultdict(lambda: [0, 0.0])
for _, name, g, b, us in step:
    short = name.split("(")[0]
    tot[short][0] += 1
    tot[short][1] += us
all_us = sum(v[1] for v in tot.values())
print(f"last step: {len(step)} launches, {all_us / 1e3:.3f} ms summed device time (cold-cache, serialised)")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us / 1e3:9.3f} ms  {100 * us / all_us:5.1f}%  x{c:5d}  {k}")
print("conv launches of the step (in order):")
for i, name, g, b, us in step:
    if "conv1d" in name or "conv2d" in name or "rvq" in name or "igemm" in name:
        print(f"  id {i:6d} {us:10.1f} us grid {g:>18s} {name.split('(')[0][-60:]}")
