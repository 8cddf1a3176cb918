"""`python -m titsolver_b200` — the reference's `titwcsph` run from Python.

Same case, loop and output as /root/reference/source/titwcsph/wcsph.cpp:157-193
(and examples/dam_break_2d.cpp / dam_break_3d.cpp): SSPRK3 steps of the dam
break until t sqrt(g / H) = 10, all particle fields as one frame of
`particles.ttdb` at the start and every 100 steps, only the last run kept.
Runs on the GPU through the C ABI; there is no CPU path.

    python -m titsolver_b200 [--dim 2] [--n-col 80] [--end-time 10] [--max-steps 0]
                             [--frame-every 100] [--out particles.ttdb] [--xdmf DIR]
"""
from __future__ import annotations

import argparse
import math
import sys
import time as _time

from . import Solver, cases, load_case, ttdb, xdmf

KERNELS = ("cubic", "quartic", "quintic", "wendland4", "wendland6", "wendland8")  # TITGPU_KERNEL_* ids 0..5
INTEGRATORS = ("euler", "verlet", "ssprk2", "ssprk3")  # TITGPU_* integrator ids 0..3


def parse(argv=None):
    p = argparse.ArgumentParser(prog="python -m titsolver_b200", description=__doc__.split("\n\n")[0])
    p.add_argument("--dim", type=int, choices=(2, 3), default=2)
    p.add_argument("--n-col", type=int, default=80, help="particles per column height H (the reference: 80)")
    p.add_argument("--end-time", type=float, default=10.0, help="in units of sqrt(H / g)")
    p.add_argument("--max-steps", type=int, default=0, help="0 = run to --end-time")
    p.add_argument("--frame-every", type=int, default=100)
    p.add_argument("--out", default="particles.ttdb", help="'-' = no database")
    p.add_argument("--xdmf", default=None, metavar="DIR", help="also export the run for ParaView into DIR")
    p.add_argument("--kernel", choices=KERNELS, default="wendland6")
    p.add_argument("--integrator", choices=INTEGRATORS, default="ssprk3")
    p.add_argument("--device", type=int, default=0)
    a = p.parse_args(argv)
    if a.n_col < 2 or a.frame_every < 1 or a.max_steps < 0:
        p.error("--n-col >= 2, --frame-every >= 1, --max-steps >= 0")
    if a.xdmf is not None and a.out == "-":
        p.error("--xdmf needs a database (--out)")
    return a


def main(argv=None) -> int:
    a = parse(argv)
    case = cases.dam_break_2d(a.n_col) if a.dim == 2 else cases.dam_break_3d(a.n_col)
    solver = Solver(a.dim, KERNELS.index(a.kernel), 0, INTEGRATORS.index(a.integrator), device=a.device)
    load_case(solver, case)
    solver.initialize()
    storage = series = None
    if a.out != "-":
        storage = ttdb.Storage(a.out)
        storage.set_max_series(1)
        series = storage.create_series()
        ttdb.write_solver_frame(series, 0.0, solver)
    scale = math.sqrt(case.g / case.H)
    t, step, started = 0.0, 1, _time.perf_counter()
    while True:
        scaled = t * scale
        last = scaled >= a.end_time or (a.max_steps and step >= a.max_steps)
        frame = step % a.frame_every == 0 or last
        solver.set_outputs(2 if frame else 0)  # derived fields only where a frame follows
        dt = solver.step(1)
        if frame:
            print(f"{step:>15}\t\t{scaled:>10.5f}\t\t{(_time.perf_counter() - started) / step:>10.5f} s/step\t\tdt = {dt:.6e}", flush=True)
            if series is not None and scaled > series.last_frame().time:
                ttdb.write_solver_frame(series, scaled, solver)
        if last:
            break
        t += dt
        step += 1
    print(f"steps {step}  particles {case.n} ({case.n_fluid} fluid)")
    if a.xdmf is not None:
        print("exported", xdmf.export_xdmf(a.xdmf, series))
    if storage is not None:
        storage.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
