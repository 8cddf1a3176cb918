#!/usr/bin/env python3
"""Shared-memory-staged pair passes (csrc/tile.cuh) against the gather traversal on the same
inputs: TITGPU_TILES=0/1 in one process. Prints per-field differences and kernel times.
    python tools/tile_check.py [n_col] [steps] [--lattice]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import titsolver_b200 as tb
from titsolver_b200 import cases


def run(case, tiles, steps):
    g = tb.Solver(3)
    g.set_tiles(tiles)
    tb.load_case(g, case)
    g.initialize()
    g.rhs_only()
    out = {f: g.download(f) for f in ("drho_dt", "dv_dt", "gamma")}
    g.profile(True)
    g.profile_reset()
    g.step(steps)
    prof = g.profile_read()
    out.update({f: g.download(f) for f in ("r", "v", "rho", "N", "phi", "grad_rho")})
    return out, prof


def main():
    n_col = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    case = cases.dam_break_3d(n_col) if "--lattice" in sys.argv else cases.dam_break_3d(n_col, wall_ratio=0.93, jitter=0.1)
    a, pa = run(case, True, steps)
    b, pb = run(case, False, steps)
    nf = case.n_fluid
    res = {"n": case.n, "steps": steps}
    for f in a:
        x, y = a[f][:nf], b[f][:nf]
        res["diff_" + f] = float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
    res["ms_tiles"] = {k: round(v[1] / max(v[0], 1), 4) for k, v in pa.items() if "rhs" in k or "shift" in k or "tile" in k}
    res["ms_gather"] = {k: round(v[1] / max(v[0], 1), 4) for k, v in pb.items() if "rhs" in k or "shift" in k}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
