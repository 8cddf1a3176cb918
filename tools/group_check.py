#!/usr/bin/env python3
"""Grouped sweep (k_rhs_grp, TITGPU_GROUP_SWEEP=1) against the gather traversal on the same inputs.
    python tools/group_check.py [dim] [n_col] [steps] [--lattice]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import titsolver_b200 as tb
from titsolver_b200 import cases


def run(case, grouped, steps):
    g = tb.Solver(case.dim)
    g.set_group_sweep(1 if grouped else 0)
    g.set_graphs(False)
    tb.load_case(g, case)
    g.initialize()
    g.rhs_only()
    out = {f: g.download(f) for f in ("drho_dt", "dv_dt", "gamma")}
    g.step(2)
    g.profile(True)
    g.profile_reset()
    g.step(steps)
    prof = g.profile_read()
    out.update({f: g.download(f) for f in ("r", "v", "rho", "N", "phi")})
    return out, prof


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n_col = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    if dim == 2:
        case = cases.dam_break_2d(n_col)
    else:
        case = cases.dam_break_3d(n_col) if "--lattice" in sys.argv else cases.dam_break_3d(n_col, wall_ratio=0.93, jitter=0.1)
    a, pa = run(case, True, steps)
    b, pb = run(case, False, steps)
    nf = case.n_fluid
    res = {"dim": dim, "n": case.n, "steps": steps}
    for f in a:
        x, y = a[f][:nf], b[f][:nf]
        res["diff_" + f] = float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
    res["ms_grouped"] = {k: round(v[1] / max(v[0], 1), 4) for k, v in pa.items() if "rhs" in k or "shift" in k}
    res["ms_gather"] = {k: round(v[1] / max(v[0], 1), 4) for k, v in pb.items() if "rhs" in k or "shift" in k}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
