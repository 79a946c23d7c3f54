#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libwarpsense_b200.so the instruction count and the opcodes that show how
data moves (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, REDG = fire-and-forget global reduction,
ATOMG = global atomic with return, warp collectives).  Usage: python tools/sass_evidence.py > profiles/<tag>_sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "warpsense_b200", "libwarpsense_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = ("UBLKCP", "SYNCS", "LDGSTS", "REDG", "ATOMG", "ATOMS", "REDUX", "VOTE", "SHFL", "MATCH", "MUFU", "DFMA", "DMUL", "DADD",
        "IMAD", "LDG", "STG", "LDS", "STS", "BAR", "HMMA", "UTCMMA", "UTMALDG", "MEMBAR", "ERRBAR")
print("# SASS evidence (cuobjdump -sass %s, sm_100a)" % os.path.relpath(lib, ROOT))
print("# Per kernel: instruction count and the opcodes that show how data moves: UBLKCP = cp.async.bulk (TMA bulk copy),")
print("# SYNCS = mbarrier, LDGSTS = cp.async (global -> shared), REDG = fire-and-forget global reduction, ATOMG = global atomic")
print("# with return, REDUX/VOTE/SHFL/MATCH = warp collectives.  No HMMA/UTCMMA: there is no dense contraction on this path.\n")
cur, counts, n = None, None, 0


def flush():
    if cur is None:
        return
    items = ", ".join("%s x%d" % (k, v) for k, v in sorted(counts.items()))
    print(cur[:150])
    print("    %d instructions; %s" % (n, items))


for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        flush()
        cur, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur is not None:
        n += 1
        op = m.group(1)
        base = op.split(".")[0]
        if base in KEEP:
            key = op if base in ("REDG", "ATOMG", "ATOMS", "LDGSTS", "UBLKCP", "SYNCS") else base
            counts[key] += 1
flush()
