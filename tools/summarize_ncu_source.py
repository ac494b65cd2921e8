"""Top source lines by warp-stall samples from `ncu --page source --csv` (needs -lineinfo + --import-source on)."""
import csv
import sys
from collections import defaultdict

with open(sys.argv[1], newline="") as f:
    rows = list(csv.reader(f))
# the csv holds one table per kernel: header rows start with "#" or contain "Source"
hdr = None
agg = defaultdict(lambda: [0.0, 0.0])
kernel = "?"
for r in rows:
    if not r:
        continue
    if "Source" in r and any("Samples" in c for c in r):
        hdr = {c: i for i, c in enumerate(r)}
        continue
    if len(r) == 1 or (hdr is None):
        if r and "Kernel" in r[0]:
            kernel = r[0][:80]
        continue
    try:
        src = r[hdr["Source"]].strip()
        sm = [c for c in hdr if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
        val = 0.0
        for c in hdr:
            if "Warp Stall Sampling (All" in c:
                val = float(r[hdr[c]].replace(",", "") or 0)
        if val:
            agg[(kernel, src[:110])][0] += val
    except Exception:
        continue
tot = sum(v[0] for v in agg.values()) or 1.0

for (k, s), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:55]:
    print(f"{100 * v[0] / tot:5.1f}%  {s}")
