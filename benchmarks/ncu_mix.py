#!/usr/bin/env python
"""Dynamic instruction mix and stall samples per opcode from an `ncu --page source --csv` export.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv ; python benchmarks/ncu_mix.py src.csv"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, sm = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[iS])
    if not m:
        continue
    op = m.group(2)
    ex[op] += int(r[iE] or 0)
    sm[op] += int(r[iN] or 0)
tot, tots = sum(ex.values()), sum(sm.values())
print(f"total warp instructions {tot}, samples {tots}")
for op, c in ex.most_common(18):
    print(f"{op:10s} {c:12d} {100*c/tot:5.1f}%   samples {100*sm[op]/max(tots,1):5.1f}%")
