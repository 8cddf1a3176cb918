"""The drop-in claim, proven in the tree: the reference's own driver
(/root/reference/source/titwcsph/wcsph.cpp, UNMODIFIED apart from the include swap of
INTEGRATION.md section 1) compiles against include/tit_b200/sph.hpp, links with
libtitgpu.so, runs on the GPU and reproduces the oracle.

The source is read where it lies at build time (`__graft_entry__.build_reference_driver`);
the GPU box has no /root/reference and runs the prebuilt, git-ignored binary."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from titsolver_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "wcsph_reference_driver")


@pytest.mark.skipif(not os.path.exists(entry.REFERENCE_DRIVER), reason="the reference tree is not present on this machine")
def test_reference_driver_compiles_against_the_facade():
    if os.path.exists(EXE):
        os.remove(EXE)
    exe = entry.build_reference_driver()
    assert exe == EXE and os.path.exists(exe)
    # the shim supplies exactly the four out-of-scope core names and stays small
    with open(os.path.join(ROOT, "tests", "cpp", "wcsph_shim.hpp")) as f:
        code = [l for l in f if l.strip() and not l.lstrip().startswith("//")]
    assert len(code) <= 36


@pytest.mark.gpu
def test_reference_driver_runs_on_the_gpu_and_matches_the_oracle(tmp_path, oracle):
    if not os.path.exists(EXE):
        pytest.skip("examples/wcsph_reference_driver was not built (needs /root/reference at build time)")
    from titsolver_b200 import ttdb

    steps = 100  # the driver writes a frame at t = 0 and after every 100th step (wcsph.cpp:184-189)
    env = dict(os.environ, TIT_SHIM_MAX_STEPS=str(steps))
    r = subprocess.run([EXE], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.splitlines()) == steps
    with ttdb.Storage(str(tmp_path / "particles.ttdb"), read_only=True) as st:
        series = st.series()[-1]
        assert series.num_frames == 2
        first, last = series.frame(0).read(), series.last_frame().read()
    case = cases.dam_break_2d(80)  # the driver's own constants (wcsph.cpp:37-142)
    nf = case.n_fluid
    assert first["r"].shape == (case.n, 2) and np.array_equal(first["r"], case.r)
    assert np.allclose(first["rho"][:nf], case.rho[:nf], rtol=1e-15, atol=0)
    cpu = oracle.OracleSolver(2)
    oracle.load_case(cpu, case)
    cpu.initialize()
    for _ in range(steps):
        cpu.step(1)
    for f, tol in (("r", 1e-9), ("v", 1e-5), ("rho", 1e-8)):
        a, b = last[f][:nf], cpu.download(f)[:nf]
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        assert err <= tol, (f, err)
    # the observable fields of the last step were published too
    for f in ("gamma", "p", "dv_dt", "N", "phi"):
        a, b = last[f][:nf], cpu.download(f)[:nf]
        assert np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) <= 1e-4, f
