#!/usr/bin/env python3
"""Milliseconds per step of the launch-bound sizes, with whole-step CUDA graphs on and off.
    python tools/small_case_timing.py [n_col ...]      (2-D dam break; 80 = the reference's default case C1)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import titsolver_b200 as tb
from titsolver_b200 import cases

for n_col in [int(x) for x in sys.argv[1:]] or [80, 200, 707]:
    case = cases.dam_break_2d(n_col)
    out = {"case": case.meta["name"], "n": case.n}
    for graphs in (True, False):
        g = tb.Solver(2)
        g.set_graphs(graphs)
        g.set_outputs(0)
        tb.load_case(g, case)
        g.initialize()
        g.step(20)
        steps = 400 if case.n < 200000 else 50
        g.synchronize()
        t0 = time.perf_counter()
        g.step(steps)
        g.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        out["ms_per_step_graphs_on" if graphs else "ms_per_step_graphs_off"] = round(ms, 4)
        out["updates_per_s_graphs_on" if graphs else "updates_per_s_graphs_off"] = round(case.n / ms * 1e3)
        if graphs:
            out["graph_replays"] = g.graph_replays
    print(json.dumps(out))
