#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (a --set full capture): one block
of the roofline-relevant metrics per profiled launch.
usage: ncu_full_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("gpu__time_duration.sum", "device time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts % of peak"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "L1 wavefronts, global loads"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "L1 wavefronts, local (spill) loads"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "L1 wavefronts, local (spill) stores"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "L1 wavefronts, shared"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "FMA (FP32) pipe %"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-instr"),
]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("== " + r[idx["Kernel Name"]].replace("void ", "")[:70])
    for m, label in WANT:
        if m in idx:
            v = r[idx[m]]
            try:
                v = f"{float(v):,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"  {label:42s} {v:>22s} {units[idx[m]]}")
    rd, wr, t = (float(r[idx[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    ur, ut = units[idx["dram__bytes_read.sum"]], units[idx["gpu__time_duration.sum"]]
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[ur]
    st = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(ut, 1e-3)
    scw = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[idx["dram__bytes_write.sum"]]]
    print(f"  {'DRAM traffic / launch, achieved GB/s':42s} {(rd*sc+wr*scw)/1e9:>14.3f} GB {(rd*sc+wr*scw)/1e9/(t*st):>10.1f} GB/s")
