#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list.
usage: ncu_launch_summary.py launches.csv "<header note>" > profiles/<name>.txt"""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[idx["Kernel Name"]].replace("void titgpu::", "").replace("titgpu::", "").split("(")[0]
    v = float(r[idx["Metric Value"]].replace(",", "")); u = r[idx["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print()
print(f"{'kernel':42s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:42]:42s} {c:8d} {ms:10.3f} {ms / tot * 100:6.1f}%")
print(f"{'total':42s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")
