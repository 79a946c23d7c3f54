#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = {}
cur_file = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    ln = r[0]
    if not ln.strip(): continue
    iex = hdr.index("Instructions Executed"); smp = hdr.index("# Samples")
    lsb = hdr.index("stall_long_sb")
    try:
        key = (cur_file.split("/")[-1], int(ln))
    except ValueError:
        continue
    a = agg.setdefault(key, [r[1], 0, 0, 0])
    if r[2] == "-":   # per-line summary row (the SASS rows that follow have an address here)
        a[1] += int(r[iex] or 0); a[2] += int(r[smp] or 0); a[3] += int(r[lsb] or 0)
tot = sum(v[1] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
top = sorted(agg.items(), key=lambda kv: -kv[1][2])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
print("total inst %d samples %d" % (tot, tots))
for (f, ln), v in top:
    print("%-18s %4d  inst %5.1f%%  samples %5.1f%%  long_sb %6d  | %s" % (f, ln, 100.0 * v[1] / tot, 100.0 * v[2] / tots, v[3], v[0].strip()[:110]))
