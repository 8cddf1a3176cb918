#!/usr/bin/env python3
"""Kernel times of repeated stand-alone RHS evaluations (state unchanged) for the
library selected by TITGPU_LIB: timing experiments on k_rhs variants.
usage: rhs_only_times.py [dim] [n_col] [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import titsolver_b200 as tb
from titsolver_b200 import cases
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_col = int(sys.argv[2]) if len(sys.argv) > 2 else 110
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
case = cases.dam_break_2d(n_col) if dim == 2 else cases.dam_break_3d(n_col)
s = tb.Solver(dim)
tb.load_case(s, case)
s.initialize()
s.rhs_only()
s.synchronize()
s.profile(True)
s.profile_reset()
for _ in range(reps):
    s.rhs_only()
s.synchronize()
prof = s.profile_read()
print(json.dumps({"lib": os.path.basename(tb.LIB_PATH), "n": case.n, "ms_per_call": {k: round(v[1] / reps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]}}))
