#!/usr/bin/env python3
"""Opcode histogram of an `ncu -i rep --page source --csv` dump: executed warp instructions and stall samples per opcode.
usage: ncu -i x.ncu-rep --page source --csv > src.csv; tools/ncu_sass_hist.py src.csv [top]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iX, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); smp = collections.Counter()
tot = 0; tots = 0
for r in rows[2:]:
    if len(r) <= iX: continue
    src = r[iS].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LDGSTS", "SYNCS", "BAR", "SHFL", "LDTM", "STTM") else op.split(".")[0]
    n = int(r[iX] or 0); s = int(r[iN] or 0)
    ops[op] += n; smp[op] += s; tot += n; tots += s
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"total warp instructions {tot:.4g}, samples {tots}")
for op, n in ops.most_common(top):
    print(f"{op:22s} {n:14d} {100*n/tot:6.2f}%   samples {smp[op]:8d} {100*smp[op]/max(1,tots):6.2f}%")
