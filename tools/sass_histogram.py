"""Per-kernel SASS opcode histogram of the built library (evidence that the hot kernels really use tcgen05 / TMEM / TMA):
   python tools/sass_histogram.py > profiles/sass_opcodes_<round>.txt
Counts the Blackwell-specific opcodes (B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc), UBLKCP (cp.async.bulk), UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "funcodec_b200", "lib", "libfuncodec_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA2", "FFMA",
        "F2FP", "MUFU.EX2", "LDGSTS", "REDG", "RED.", "LDL", "STL"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        hist[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[cur][k] += 1
                break
digest = subprocess.run(["sha256sum", lib], capture_output=True, text=True).stdout.split()[0][:16]
print(f"# SASS opcode histogram of funcodec_b200/lib/libfuncodec_b200.so (sha256 {digest}...), sm_100a")
print("# kernel".ljust(64) + "".join(k.rjust(11) for k in ["total"] + KEYS))
for name, c in hist.items():
    if c["total"] < 50:
        continue
    print(name[:63].ljust(64) + "".join(str(c[k]).rjust(11) for k in ["total"] + KEYS))
