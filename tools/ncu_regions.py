#!/usr/bin/env python3
"""tools/ncu_regions.py <source.csv> — splits the `ncu --page source --csv` listing of a gen3 resampler kernel into the code before
the tap walk, the walk (first .. last hot FFMA2) and the code after it: stall-sample share, executed-instruction share, opcode
counts per warp and group, stall reasons, hottest non-walk instructions (development aid)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ins.append((int(r[col['Address']], 16), r[col['Source']], int(r[col['# Samples']] or 0), int(r[col['Instructions Executed']] or 0), r))
    except ValueError:
        pass
tot_s = sum(i[2] for i in ins)
tot_e = sum(i[3] for i in ins)
per_group = int(sys.argv[2]) if len(sys.argv) > 2 else 2112
groups = sum(i[3] for i in ins if 'FFMA2' in i[1]) / per_group
hot = [k for k, i in enumerate(ins) if 'FFMA2' in i[1] and i[3] > groups * 0.9 / 12]
first, last = min(hot), max(hot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"{len(ins)} instructions, walk = [{first}, {last}], {groups:.0f} warp-groups")
for name, (a, b) in {'before': (0, first), 'walk': (first, last + 1), 'after': (last + 1, len(ins))}.items():
    seg = ins[a:b]
    d = defaultdict(float)
    st = defaultdict(int)
    for i in seg:
        op = i[1].split()[0] if not i[1].startswith('@') else i[1].split()[1]
        d[op] += i[3] / groups
        for h in stalls:
            st[h[6:]] += int(i[4][col[h]] or 0)
    print(f"{name}: samples {100 * sum(i[2] for i in seg) / tot_s:.1f}%  executed {100 * sum(i[3] for i in seg) / tot_e:.1f}%  per group {sum(d.values()):.0f}")
    print("    ops:", ", ".join(f"{k} {v:.0f}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:14]))
    print("    stalls:", ", ".join(f"{k} {100 * v / tot_s:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
print("hottest instructions outside the walk:")
n = 0
for k, i in sorted(enumerate(ins), key=lambda x: -x[1][2]):
    if first <= k <= last:
        continue
    st = sorted(((int(i[4][col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  {k:5d} {100 * i[2] / tot_s:4.2f}% x{i[3] / groups:4.2f} {i[1][:64]:64s}", st)
    n += 1
    if n >= (int(sys.argv[3]) if len(sys.argv) > 3 else 16):
        break
