"""Slab decomposition on the GPU: N ranks (one context each) must reproduce the
single-context run of the same case. On a single-GPU box the ranks are threads of
one process sharing the device (`--hub N`: event-ordered device copies instead of
NCCL, the same exchange kernels - csrc/mg.cuh); with one GPU per rank
tools/slab_check.py runs under torchrun over NCCL (profiles/). Results depend on
the decomposition only through the order of the floating-point sums, hence the
1e-9 bound after a few steps."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_check(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "slab_check.py"), *map(str, args)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    res = json.loads(lines[-1])
    assert res["ok"], res
    return res


def test_two_slabs_2d():
    # 80 particle layers along x: each slab (40) is wider than the halo (~10 dr)
    res = run_check("--hub", 2, "--dim", 2, "--n-col", 40, "--steps", 3)
    assert all(c[1] > 0 for c in res["counts"])  # both ranks hold ghosts
    assert all(e[0] >= 3 * 6 for e in res["exchanges_migrated"])  # per step: halo set, 3 refreshes, N / phi, shifted records


def test_four_slabs_2d_with_migration():
    """A random velocity kick makes particles change owner during the run; interior
    slabs have two neighbours."""
    res = run_check("--hub", 4, "--dim", 2, "--n-col", 40, "--steps", 6, "--kick", 20.0, "--tol", 1e-7)
    assert sum(e[1] for e in res["exchanges_migrated"]) > 0  # particles did migrate


def test_two_slabs_3d():
    run_check("--hub", 2, "--dim", 3, "--n-col", 24, "--steps", 2)


def test_three_slabs_3d_benchmark_lattice():
    run_check("--hub", 3, "--dim", 3, "--n-col", 20, "--steps", 2, "--lattice")


def test_two_slabs_3d_along_z():
    run_check("--hub", 2, "--dim", 3, "--n-col", 12, "--steps", 2, "--axis", 2)


def test_two_slabs_2d_euler_and_verlet():
    for integ in (0, 1):
        run_check("--hub", 2, "--dim", 2, "--n-col", 30, "--steps", 3, "--integrator", integ)


def test_slabs_with_the_grouped_candidate_sweep():
    """The grouped sweep (k_rhs_grp / k_shift_grp, forced on here: it is chosen by size otherwise)
    skips ghosts as members but reads them as neighbours, like the default traversal."""
    run_check("--hub", 2, "--dim", 3, "--n-col", 24, "--steps", 2, env={"TITGPU_GROUP_SWEEP": "1"})
    run_check("--hub", 3, "--dim", 2, "--n-col", 60, "--steps", 4, "--kick", 20.0, "--tol", 1e-7, env={"TITGPU_GROUP_SWEEP": "1"})


def test_slabs_with_the_tiled_cell_order():
    """The y-tiled cell order (chosen by the grid on large cross-sections, forced to tiles of 4 cells here)
    under a slab decomposition, with the grouped sweep on: ghosts, migration and the exchanges do not
    depend on the order of the sorted arrays."""
    env = {"TITGPU_YTILE_LOG": "2", "TITGPU_GROUP_SWEEP": "1"}
    run_check("--hub", 2, "--dim", 3, "--n-col", 24, "--steps", 3, env=env)
    run_check("--hub", 3, "--dim", 3, "--n-col", 20, "--steps", 2, "--lattice", env=env)
