#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: hot SASS regions by executed
instructions and stall samples. usage: ncu_regions.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["Instructions Executed"]].isdigit()]
ie = idx['Instructions Executed']; ss = idx['Warp Stall Sampling (All Samples)']; te = idx['Thread Instructions Executed']
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[ss]) for r in data)
print('total warp inst', tot, 'SASS rows', len(data), 'stall samples', tots)
i = 0
while i < len(data):
    c = int(data[i][ie]); j = i; s = 0; t = 0
    while j < len(data) and abs(int(data[j][ie]) - c) <= 0.02 * max(c, 1):
        s += int(data[j][ss]); t += int(data[j][te]); j += 1
    if c * (j - i) > minp / 100 * tot or s > minp / 100 * tots:
        ops = {}
        for r in data[i:j]:
            op = r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = ' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"rows {i:5d}-{j:5d} n={j-i:4d} exec={c:10d} inst%={100*c*(j-i)/tot:5.1f} stall%={100*s/max(tots,1):5.1f} thr/inst={t/max(c*(j-i),1):4.1f} | {top}")
    i = j
