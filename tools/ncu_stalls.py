#!/usr/bin/env python3
"""Aggregate the per-instruction stall samples of an `ncu --page source --csv` dump.
usage: ncu_stalls.py <source.csv> [lo_off hi_off]   (offsets relative to the kernel start, hex)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
tot = collections.Counter()
byop = collections.defaultdict(collections.Counter)
nsamp = 0
allsamp = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    s = int(r[ix["# Samples"]] or 0)
    allsamp += s
    if not (lo <= off <= hi):
        continue
    nsamp += s
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        tot[c] += v
        byop[op][c] += v
    byop[op]["_n"] += s
    byop[op]["_exec"] += int(r[ix["Instructions Executed"]] or 0)
print(f"samples in range {nsamp} of {allsamp}")
print("stall totals:", [(k, v) for k, v in tot.most_common(8)])
for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["_n"])[:14]:
    top = [(k, v) for k, v in c.most_common(6) if not k.startswith("_")][:4]
    print(f"{op:10s} samples {c['_n']:7d} exec {c['_exec']:10d}  {top}")
