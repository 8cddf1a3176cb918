#!/usr/bin/env python3
"""Profiling driver: one SSPRK3 step (no derived-field output) inside a
cudaProfilerStart/Stop range, for `ncu --profile-from-start off`.
usage: prof_step.py [dim] [n_col] [steps_in_range]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import titsolver_b200 as tb
from titsolver_b200 import cases

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_col = int(sys.argv[2]) if len(sys.argv) > 2 else 100
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
case = cases.dam_break_2d(n_col) if dim == 2 else cases.dam_break_3d(n_col)
s = tb.Solver(dim)
tb.load_case(s, case)
s.initialize()
s.step(3)
s.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
s.step(k + 1)  # the last step of a call also publishes the derived fields
s.synchronize()
rt.cudaProfilerStop()
print("n", case.n, "n_fluid", case.n_fluid)
