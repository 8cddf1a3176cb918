#!/usr/bin/env python3
"""Per-kernel device time of the step for the library selected by TITGPU_LIB
(tuning variants built with TITGPU_VARIANT / TITGPU_DEFINES, see build.py), and a
checksum of the state so that variants can be compared for equal results.
usage: variant_times.py [dim] [n_col] [steps] [warmup]"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import titsolver_b200 as tb
from titsolver_b200 import cases

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_col = int(sys.argv[2]) if len(sys.argv) > 2 else 110
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 3
case = cases.dam_break_2d(n_col) if dim == 2 else cases.dam_break_3d(n_col)
s = tb.Solver(dim)
tb.load_case(s, case)
s.set_lists(int(os.environ.get('TIT_LISTS', '0')))
s.initialize()
s.set_outputs(0)
s.step(warm)
s.synchronize()
s.profile(True)
s.profile_reset()
s.step(steps)
s.synchronize()
prof = s.profile_read()
s.profile(False)
h = hashlib.sha1()
sums = {}
for f in ("r", "v", "rho"):
    a = s.download(f)
    h.update(np.ascontiguousarray(a).tobytes())
    sums[f] = float(np.abs(a).sum())
tot = sum(v[1] for v in prof.values())
out = {"lib": os.path.basename(tb.LIB_PATH), "n": case.n, "ms_per_step": round(tot / steps, 3),
       "kernels_ms_per_step": {k: round(v[1] / steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:9]},
       "lists": int(os.environ.get("TIT_LISTS", "0")), "redos": s.list_redos, "sha1": h.hexdigest()[:12], "abs_sums": sums}
print(json.dumps(out), flush=True)
