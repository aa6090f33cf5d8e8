#!/usr/bin/env python
"""Hot SASS regions (runs of instructions with the same execution count) from an `ncu --page source --csv` export.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv ; python benchmarks/ncu_regions.py src.csv [min_share_pct]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
hdr = rows[1]; iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iN = hdr.index("# Samples")
data = [(int(r[iE] or 0), r[iS].strip(), int(r[iN] or 0)) for r in rows[2:] if len(r) > iE]
regions, cur = [], None
for c, s, n in data:
    if cur and cur[0] == c: cur[1].append(s); cur[2] += n
    else:
        cur = [c, [s], n]; regions.append(cur)
tot = sum(c for c, _, _ in data); tots = sum(n for _, _, n in data)
for c, ins, n in regions:
    if c * len(ins) > thr / 100 * tot:
        ops = collections.Counter(re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", x).group(2) for x in ins)
        print(f"count/instr={c} n={len(ins)} instr-share={100*c*len(ins)/tot:.1f}% sample-share={100*n/max(tots,1):.1f}%", dict(ops.most_common(9)))
