"""The C++ facade (include/tit_b200/sph.hpp): the reference driver restated
against it compiles and links on CPU; on the GPU it reproduces the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_example():
    import __graft_entry__ as ge

    return ge.build_examples()


def test_facade_driver_compiles_and_links():
    exe = build_example()
    assert os.access(exe, os.X_OK)
    # every titgpu_* symbol the facade calls is an undefined reference resolved by libtitgpu.so
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    used = {line.split()[-1] for line in out.splitlines() if "titgpu_" in line}
    assert {"titgpu_create", "titgpu_set_params", "titgpu_set_surface", "titgpu_upload", "titgpu_download", "titgpu_initialize", "titgpu_step", "titgpu_neighbors"} <= used


@pytest.mark.parametrize("std", ["c++20", "c++23"])
def test_facade_header_standalone(std, tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('''#include "tit_b200/sph.hpp"
int main() {
  using namespace tit;
  // the alternative index / partition options of the reference are accepted (SURVEY §8a-a5, a7)
  sph::ParticleMesh kd{geom::KDTreeSearch{}, geom::GridFaceSearch{0.1}, geom::RecursiveCoordBisection{}, geom::KMeansClustering{1e-3, 5}};
  sph::ParticleMesh sorted{geom::kd_tree_indexing, geom::GridFaceSearch{0.1}, geom::hilbert_curve_partition, geom::PixelatedPartition{0.2, geom::kmeans_clustering}};
  if (kd.search_hint() != 0.0 || sorted.face_search_hint() != 0.1) return 1;
  Vec v{1.0, 2.0};
  return int(dot(v, v)) - 5;
}
''')
    subprocess.check_call(["/usr/bin/g++", f"-std={std}", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", str(src)])


@pytest.mark.gpu
def test_facade_driver_matches_oracle(oracle, tmp_path):
    """examples/dam_break_2d.cpp (the reference's default case through the facade) for 5 steps
    of the 2-D dam break vs the oracle fed by the Python case generator."""
    from titsolver_b200 import cases

    exe = build_example()
    dump = tmp_path / "dump.bin"
    r = subprocess.run([exe, "20", "5", str(dump)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    case = cases.dam_break_2d(20)
    raw = np.fromfile(dump)
    n = case.n
    assert raw.size == 5 * n
    got = {"r": raw[: 2 * n].reshape(n, 2), "v": raw[2 * n: 4 * n].reshape(n, 2), "rho": raw[4 * n:]}
    c = oracle.OracleSolver(2)
    oracle.load_case(c, case)
    c.initialize()
    c.step(5)
    for f, tol in (("r", 1e-10), ("v", 1e-8), ("rho", 1e-10)):
        ref = c.download(f)
        assert np.abs(got[f] - ref).max() <= tol * np.abs(ref).max(), f
    assert "|mesh[0]| = " in r.stdout


@pytest.mark.gpu
def test_facade_reports_errors_as_exceptions(tmp_path):
    src = tmp_path / "e.cpp"
    src.write_text('''#include <cstdio>
#include "tit_b200/sph.hpp"
using namespace tit; using namespace tit::sph;
int main() {
  geom::Surface<Vec<double, 2>> s;
  const auto w = geom::make_exact_winding(s);
  const FluidEquations eq{9.81, 1e-3, s, w, TaitEquationOfState{10.0, 1000.0}, CubicSplineKernel{}};
  const SymplecticEulerIntegrator ti{eq};
  const SSPRKIntegrator other{eq};
  ParticleArray particles{Space<double, 2>{}, ti};
  ParticleMesh mesh{geom::GridSearch{0.1}, geom::GridFaceSearch{0.1}};
  try { other.step(mesh, particles); } catch (const Exception& e) { std::printf("caught: %s\\n", e.what()); return 0; }
  return 1;
}
''')
    exe = tmp_path / "e"
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{ROOT}/titsolver_b200", "-ltitgpu", f"-Wl,-rpath,{ROOT}/titsolver_b200"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "caught: " in r.stdout, (r.stdout, r.stderr)


def test_facade_2d_driver_setup_matches_python_case(tmp_path):
    """examples/dam_break_2d.cpp sets up the reference's default case exactly as
    titsolver_b200.cases.dam_break_2d does (the case the oracle and the GPU tests use)."""
    from titsolver_b200 import cases, ttdb

    db = tmp_path / "setup.ttdb"
    out = subprocess.run([build_example(), "-20", "0", "-", str(db)], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    with ttdb.Storage(str(db), read_only=True) as s:
        d = s.last_series().last_frame().read()
    c = cases.dam_break_2d(20)
    assert f"{c.n_fluid} fluid + {c.n_fixed} fixed" in out.stdout
    for got, want in (("verts", c.verts), ("faces", c.faces), ("containment_verts", c.cverts), ("containment_faces", c.cfaces), ("r", c.r), ("m", c.m), ("rho", c.rho)):
        assert np.array_equal(d[got], want), got
