"""Per-launch durations of the last bench step for several FCB_TC_DBG knock-out masks side by side.
usage: summarize_knockouts.py <tag> <mask> [<mask> ...]   (reads gpurun_out/launches_ko<mask>_<tag>.csv)"""
import csv
import sys

tag, masks = sys.argv[1], sys.argv[2:]
cols = {}
for m in masks:
    rows = []
    try:
        with open(f"gpurun_out/launches_ko{m}_{tag}.csv", newline="") as f:
            lines = [l for l in f if not l.startswith("==")]
    except OSError:
        continue
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r.get("Metric Unit", "ns"), 1e-3)
        rows.append((r["Kernel Name"].split("(")[0][-40:], v * scale))
    n = len(rows)
    cols[m] = rows[-(n // 4):]
if not cols:
    sys.exit("no data")
base = cols[masks[0]]
print("launch".ljust(48) + "".join(f"ko={m:>4s} " for m in cols))
tot = {m: 0.0 for m in cols}
for i, (name, _) in enumerate(base):
    if "conv1d_tc" not in name:
        continue
    line = f"{i:3d} {name:44s}"
    for m in cols:
        us = cols[m][i][1] if i < len(cols[m]) else float("nan")
        tot[m] += us
        line += f"{us:8.1f}"
    print(line)
print("sum of conv1d_tc launches (us)".ljust(48) + "".join(f"{tot[m]:8.0f}" for m in cols))
