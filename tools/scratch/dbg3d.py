import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import titsolver_b200 as tb
from titsolver_b200 import cases
import oracle_lib

case = cases.dam_break_3d(6)
g = tb.Solver(3); c = oracle_lib.OracleSolver(3)
tb.load_case(g, case); oracle_lib.load_case(c, case)
g.initialize(); c.initialize()
for f in ("gamma", "grad_gamma"):
    a, b = g.download(f), c.download(f)
    d = np.abs(a - b)
    if d.ndim > 1: d = d.max(1)
    idx = np.argsort(-d)[:12]
    print(f, "max", d.max(), "n>1e-10:", (d > 1e-10).sum())
    for i in idx:
        print("  ", i, "fixed" if i >= case.n_fluid else "fluid", case.r[i], a[i], b[i])
fo, fcols = c.face_neighbors()
i = int(np.argmax(np.abs(g.download("gamma") - c.download("gamma"))))
print("faces of worst", i, fcols[fo[i]:fo[i+1]].tolist())

# early perf data
def timeit(case, steps=3):
    s = tb.Solver(case.dim)
    tb.load_case(s, case)
    s.initialize()
    s.step(1); s.synchronize()
    t = time.time(); s.step(steps); s.synchronize(); dt = (time.time() - t) / steps
    print(case.meta["name"], "n=", case.n, "ms/step", dt * 1e3, "updates/s", case.n / dt, "launches", s.launch_count)
timeit(cases.dam_break_2d(707))
timeit(cases.dam_break_3d(48))
