"""Slab decomposition on the GPU: N ranks (one context each) must reproduce the
single-context run of the same case. With fewer GPUs than ranks the ranks share
a device and talk over gloo (payloads staged through the host); with enough
GPUs they use NCCL. Results depend on the decomposition only through the order
of the floating-point sums, hence the 1e-9 bound after a few steps."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_check(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "slab_check.py"), *map(str, args)], capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    res = json.loads(lines[-1])
    assert res["ok"], res
    return res


def test_two_slabs_2d():
    # 80 particle layers along x: each slab (40) is wider than the widest halo (~18.4 dr)
    res = run_check("--spawn", 2, "--dim", 2, "--n-col", 40, "--steps", 3)
    assert all(c[1] > 0 for c in res["counts"])  # both ranks hold ghosts


def test_three_slabs_2d_with_migration():
    """A random velocity kick makes particles change owner during the run."""
    run_check("--spawn", 3, "--dim", 2, "--n-col", 30, "--steps", 4, "--kick", 3.0, "--tol", 1e-8)


def test_two_slabs_3d():
    run_check("--spawn", 2, "--dim", 3, "--n-col", 24, "--steps", 2)


def test_two_slabs_3d_along_z():
    run_check("--spawn", 2, "--dim", 3, "--n-col", 10, "--steps", 2, "--axis", 2)
