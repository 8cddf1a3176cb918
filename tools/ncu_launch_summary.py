#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel launches, total / mean device time and share.
usage: ncu_launch_summary.py launches.csv [skip_first_n_launches]"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4].replace("void ", "").replace("titgpu::", ""))
    t = float(r[14]) * (1e-3 if r[13] == "ns" else 1.0)  # us
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows)} launches, {tot/1e3:.3f} ms total device time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'mean_us':>10s} {'share':>7s}  block grid(last)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {a[0]:8d} {a[1]/1e3:10.3f} {a[1]/a[0]:10.1f} {100*a[1]/tot:6.1f}%  {a[2]} {a[3]}")
