#!/usr/bin/env python3
"""Run the same case with two builds of the library (separate processes) and compare the states.
usage: compare_libs.py libA.so libB.so [n_col] [steps]"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import titsolver_b200 as tb
    from titsolver_b200 import cases
    n_col, steps, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    case = cases.dam_break_3d(n_col)
    s = tb.Solver(3); tb.load_case(s, case); s.initialize(); s.step(steps)
    np.savez(out, **{f: s.download(f) for f in ("r", "v", "rho", "gamma", "grad_gamma", "N", "phi")})
    sys.exit(0)
a, b = sys.argv[1], sys.argv[2]
n_col = sys.argv[3] if len(sys.argv) > 3 else "60"
steps = sys.argv[4] if len(sys.argv) > 4 else "6"
outs = []
for i, lib in enumerate((a, b, b)):
    out = f"/tmp/cmp_{i}.npz"
    subprocess.check_call([sys.executable, __file__, "--child", n_col, steps, out], env={**os.environ, "TITGPU_LIB": os.path.abspath(lib)})
    outs.append(np.load(out))
for name, x, y in (("A vs B", outs[0], outs[1]), ("B vs B again", outs[1], outs[2])):
    print(name, {f: float(np.abs(x[f] - y[f]).max() / max(np.abs(x[f]).max(), 1e-300)) for f in x.files})
