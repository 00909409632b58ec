#!/usr/bin/env python3
"""tools/ncu_source.py <source.csv> [top] — per-SASS-instruction stall samples from `ncu --page source --csv`,
grouped by opcode and listed by hottest instruction (development aid)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ins.append((r[col["Address"]], r[col["Source"]], int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0), r))
    except ValueError:
        continue
tot = sum(i[2] for i in ins) or 1
tot_inst = sum(i[3] for i in ins) or 1
byop_s, byop_i = defaultdict(int), defaultdict(int)
for a, s, n, e, r in ins:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    byop_s[op] += n
    byop_i[op] += e
print("opcode: samples% / executed%")
for op, n in sorted(byop_s.items(), key=lambda kv: -kv[1])[:16]:
    print(f"  {op:24s} {100 * n / tot:5.1f}%  {100 * byop_i[op] / tot_inst:5.1f}%")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("hottest instructions:")
for a, s, n, e, r in sorted(ins, key=lambda x: -x[2])[:top]:
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"  {a[-5:]} {100 * n / tot:4.1f}% {s[:70]:70s} " + " ".join(f"{k}:{v}" for v, k in st if v))
