#!/usr/bin/env python3
"""Record the DRAM traffic and pipe utilisation of one profiled launch in profiles/ncu_traffic.json
(what bench.py reports as roofline.traffic / roofline.pipes).
usage: ncu_traffic_update.py raw.csv <kernel substring> <2d|3d> <n particles of the launch> "<source note>" [launch index]
raw.csv = `ncu -i X.ncu-rep --page raw --csv` of an `ncu --set full --clock-control none` capture."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, frag, dim, n, note = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
which = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
hits = [r for r in rows[2:] if len(r) == len(hdr) and frag in r[idx["Kernel Name"]]]
r = hits[which]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
byt = sum(float(r[idx[k]]) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
g = lambda k: float(r[idx[k]])
pipes = {
    "fp64_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "fma_fp32_pct": g("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
    "alu_pct": g("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
    "lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "issue_slots_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "l1_data_pipe_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"),
    "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
    "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
}
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
name = r[idx["Kernel Name"]].split("<")[0].replace("void ", "").strip()
db.setdefault(name, {})[dim] = {
    "bytes_per_particle": byt / n, "pipes": pipes,
    "source": f"{note}: dram__bytes_read.sum + dram__bytes_write.sum = {byt / 1e9:.3f} GB for one {name} launch over n = {n} particles (ncu --set full --clock-control none)",
}
json.dump(db, open(path, "w"), indent=1)
print(name, dim, f"{byt / n:.1f} B/particle", pipes)
