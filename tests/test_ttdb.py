"""The `.ttdb` output format: titsolver_b200/ttdb.py and the C++
facade's tit::data::Storage (include/tit_b200/data.hpp).

Pinned by (i) the reference's own database fixture tests/_data/particles.ttdb —
its facts live in tests/golden/ttdb_fixture.json (made by
tests/golden/make_ttdb_golden.py with an independent zstd binding), including two
of the reference's compressed blobs verbatim; (ii) the cases of the reference's
tit/data/storage.test.cpp and type.test.cpp, restated for both implementations;
(iii) each implementation reading what the other wrote.
"""
import base64
import hashlib
import json
import os
import sqlite3
import subprocess

import numpy as np
import pytest

from titsolver_b200 import ttdb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ttdb_fixture.json")))
REF_FIXTURE = "/root/reference/tests/_data/particles.ttdb"


# ---------------------------------------------------------------- type ids (type.test.cpp)
def test_type_ids_known_answers():
    assert ttdb.type_id(np.float32, ttdb.RANK_MATRIX, 3) == 0x030209  # type.test.cpp "to ID"
    assert ttdb.type_id(np.float64, ttdb.RANK_VECTOR, 2) == 131338  # SURVEY §8f-1, the fixture's Vec2
    assert ttdb.type_id(np.float64) == 65546 and ttdb.type_id(np.uint64) == 65544  # fixture: rho, parinfo
    dt, rank, dim = ttdb.decode_type(0x030209)
    assert (dt.name, rank, dim) == ("float32", ttdb.RANK_MATRIX, 3)
    assert ttdb.type_name(131338) == "Vec<float64_t, 2>"
    assert ttdb.type_name(ttdb.type_id(np.int16, ttdb.RANK_MATRIX, 3)) == "Mat<int16_t, 3>"
    assert ttdb.type_name(ttdb.type_id(np.float32)) == "float32_t"
    for kind, width in zip(ttdb._KINDS, (1, 1, 2, 2, 4, 4, 8, 8, 4, 8)):
        assert np.dtype(kind).itemsize == width
    with pytest.raises(ValueError, match="Invalid"):
        ttdb.decode_type(0x1337)
    with pytest.raises(ValueError, match="Invalid data type rank: 137."):
        ttdb.type_id(np.float32, 137, 3)
    with pytest.raises(ValueError, match="Dimensionality must be positive, but is 0."):
        ttdb.type_id(np.float32, ttdb.RANK_VECTOR, 0)
    with pytest.raises(ValueError, match="Dimensionality of a scalar must be 1, but is 2."):
        ttdb.type_id(np.float32, ttdb.RANK_SCALAR, 2)


# ---------------------------------------------------------------- the reference's fixture
def test_golden_blobs_of_the_reference_decode():
    """Frames written by the reference's streaming compressor (no content size in the
    frame header) decode to the bytes an independent zstd binding produced."""
    n = 0
    for series in GOLDEN["series"]:
        for frame in series["frames"]:
            for a in frame["arrays"]:
                assert ttdb.decode_type(a["type"])  # every type id of the fixture is understood
                if "blob_b64" in a:
                    raw = ttdb._Zstd.decompress(base64.b64decode(a["blob_b64"]), a["nbytes"])
                    assert hashlib.sha1(raw).hexdigest() == a["sha1"]
                    n += 1
    assert n == 2


def _storage_from_golden_blobs(path):
    """A database holding the reference's two verbatim blobs under its schema."""
    s = ttdb.Storage(str(path))
    series = s.create_series("golden")
    frame = series.create_frame(0.0)
    want = {}
    for fr in GOLDEN["series"][0]["frames"]:
        for a in fr["arrays"]:
            if "blob_b64" in a:
                arr = frame.create_array(a["name"])
                with s._db:
                    s._db.execute("UPDATE DataArrays SET type = ?, size = ?, data = ? WHERE id = ?", (a["type"], a["size"], base64.b64decode(a["blob_b64"]), arr.id))
                want[a["name"]] = a
    s.close()
    return want


def test_reader_on_reference_blobs_in_a_database(tmp_path):
    want = _storage_from_golden_blobs(tmp_path / "g.ttdb")
    with ttdb.Storage(str(tmp_path / "g.ttdb"), read_only=True) as s:
        got = s.last_series().last_frame().read()
    assert set(got) == set(want) == {"rho", "FS"}
    for name, a in want.items():
        assert got[name].shape == (a["size"],) and hashlib.sha1(got[name].tobytes()).hexdigest() == a["sha1"]
    assert 990.0 < got["rho"].min() <= got["rho"].max() < 1010.0  # a water density field


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="the reference tree is only present in the build container")
def test_reads_the_reference_fixture_directly():
    with ttdb.Storage(REF_FIXTURE, read_only=True) as s:
        assert s.max_series == GOLDEN["max_series"] and s.num_series == len(GOLDEN["series"])
        for series, gs in zip(s.series(), GOLDEN["series"]):
            assert series.name == gs["name"] and series.num_frames == len(gs["frames"])
            for frame, gf in zip(series.frames(), gs["frames"]):
                assert frame.time == gf["time"]
                arrays = frame.arrays()
                assert [a.name for a in arrays] == [g["name"] for g in gf["arrays"]]
                for a, g in zip(arrays, gf["arrays"]):
                    assert (a.type, a.size) == (g["type"], g["size"])
                    assert hashlib.sha1(a.read().tobytes()).hexdigest() == g["sha1"]


def test_schema_matches_the_reference_fixture(tmp_path):
    with ttdb.Storage(str(tmp_path / "s.ttdb")) as s:
        tables = sorted(n for (n,) in s._db.execute("SELECT name FROM sqlite_master WHERE type = 'table' AND name NOT LIKE 'sqlite_%'"))
        assert tables == GOLDEN["tables"] == ["DataArrays", "DataFrames", "DataSeries", "Settings"]
        cols = {t: [(r[1], r[2]) for r in s._db.execute(f"PRAGMA table_info({t})")] for t in tables}
    assert cols["DataArrays"] == [("id", "INTEGER"), ("frame_id", "INTEGER"), ("name", "TEXT"), ("type", "INTEGER"), ("size", "INTEGER"), ("data", "BLOB")]
    assert cols["DataFrames"] == [("id", "INTEGER"), ("series_id", "INTEGER"), ("time", "REAL")]
    assert cols["DataSeries"] == [("id", "INTEGER"), ("name", "TEXT")]
    assert cols["Settings"] == [("id", "INTEGER"), ("max_series", "INTEGER")]


# ---------------------------------------------------------------- storage.test.cpp, Python side
def test_open_modes(tmp_path):
    assert ttdb.Storage(":memory:").path == ""
    p = tmp_path / "test.ttdb"
    ttdb.Storage(str(p)).close()
    assert p.exists()
    ttdb.Storage(str(p)).close()  # open existing
    ro = ttdb.Storage(str(p), read_only=True)
    with pytest.raises(sqlite3.OperationalError, match="attempt to write a readonly database"):
        ro.create_series("test")
    with pytest.raises(sqlite3.OperationalError, match="unable to open database file"):
        ttdb.Storage("/invalid/path/to/file.ttdb")
    with pytest.raises(sqlite3.DatabaseError, match="file is not a database"):
        ttdb.Storage(__file__, read_only=True)


def test_series_cap_and_deletion():
    s = ttdb.Storage()
    assert s.num_series == 0 and s.max_series >= 3
    a, b, c = s.create_series("1"), s.create_series("2"), s.create_series("3")
    assert [x.id for x in (a, b, c)] == [1, 2, 3] and [x.name for x in s.series()] == ["1", "2", "3"]
    assert s.last_series() == c and s.series(0) == a and s.series(2) == c
    with pytest.raises(IndexError, match="out of bounds"):
        s.series(3)
    s.set_max_series(3)
    d = s.create_series("4")  # the oldest goes
    assert s.check_series(d) and not s.check_series(a) and s.series() == [b, c, d]
    e = s.create_series("5")
    assert not s.check_series(b) and s.series() == [c, d, e]
    s.set_max_series(2)
    assert s.max_series == 2 and s.series() == [d, e]
    s.set_max_series(5)
    f = s.create_series("6")
    assert s.series() == [d, e, f]
    s.delete_series(e)
    g = s.create_series("7")
    assert g.id != e.id and s.series() == [d, f, g]  # ids are not reused


def test_frames_and_cascades():
    s = ttdb.Storage()
    series = s.create_series()
    assert series.num_frames == 0
    f1, f2, f3 = series.create_frame(0.0), series.create_frame(1.0), series.create_frame(2.0)
    assert [f.id for f in (f1, f2, f3)] == [1, 2, 3] and [f.time for f in series.frames()] == [0.0, 1.0, 2.0]
    assert series.last_frame() == f3 and series.frame(1) == f2
    with pytest.raises(IndexError, match="out of bounds"):
        series.frame(3)
    with pytest.raises(ValueError, match="greater than the last frame time"):
        series.create_frame(2.0)
    other = s.create_series()
    g = [other.create_frame(t) for t in (0.0, 1.0, 2.0)]
    assert len({f.id for f in (f1, f2, f3, *g)}) == 6
    s.delete_frame(f2)
    assert not s.check_frame(f2) and series.frames() == [f1, f3]
    assert series.create_frame(3.0).id == 7
    s.delete_series(other)
    assert not any(s.check_frame(f) for f in g) and s.check_frame(f1)


def test_arrays_roundtrip_update_delete():
    s = ttdb.Storage()
    frame = s.create_series().create_frame(0.0)
    assert frame.num_arrays == 0
    a1 = frame.create_array("array_1")
    a1.write(np.array([np.pi]))
    assert (a1.id, a1.name, a1.type, a1.size) == (1, "array_1", ttdb.type_id(np.float64), 1) and a1.read()[0] == np.pi
    a2 = frame.create_array("array_2")
    a2.write(np.array([np.e], dtype=np.float32))
    assert a2.type == ttdb.type_id(np.float32) and a2.read().dtype == np.float32 and a2.read()[0] == np.float32(np.e)
    assert frame.arrays() == [a1, a2] and frame.find_array("array_2") == a2 and frame.find_array("does_not_exist") is None
    a1.write(np.array([1.618, 3.0 ** 0.5]))  # overwrite
    assert a1.size == 2 and a1.read().tolist() == [1.618, 3.0 ** 0.5]
    rng = np.random.default_rng(123)
    vec, mat = rng.normal(size=(1000, 3)), rng.normal(size=(1000, 3, 3))
    ids = np.arange(1000, dtype=np.uint64)
    for name, val in (("v", vec), ("L", mat), ("ids", ids), ("empty", np.empty((0, 2)))):
        frame.create_array(name).write(val)
    got = frame.read()
    assert np.array_equal(got["v"], vec) and np.array_equal(got["L"], mat) and np.array_equal(got["ids"], ids) and got["empty"].shape == (0, 2)
    assert frame.find_array("L").type == ttdb.type_id(np.float64, ttdb.RANK_MATRIX, 3)
    with pytest.raises(ValueError, match="already exists"):
        frame.create_array("v")
    with pytest.raises(ValueError, match="must not be empty"):
        frame.create_array("")
    s.delete_array(a1)
    assert not s.check_array(a1) and frame.arrays()[0] == a2
    assert frame.create_array("array_3").id == 7
    s.delete_frame(frame)
    assert not s.check_array(a2)


def test_write_particles_field_order():
    n, d = 50, 3
    rng = np.random.default_rng(7)
    fields = {f: rng.normal(size=(n,)) for f in ("m", "gamma", "rho", "drho_dt", "p", "cs", "phi", "rho_raw")}
    fields.update({f: rng.normal(size=(n, d)) for f in ("grad_gamma", "grad_rho", "v", "dv_dt", "r", "dr", "N")})
    fields.update({f: rng.normal(size=(n, d, d)) for f in ("grad_v", "L")})
    s = ttdb.Storage()
    frame = s.create_series().write_particles(0.25, fields)
    assert [a.name for a in frame.arrays()] == list(ttdb.PARTICLE_FIELDS)  # fluid_equations.hpp:41-48
    got = frame.read()
    assert all(np.array_equal(got[f], fields[f]) for f in fields) and frame.time == 0.25


# ---------------------------------------------------------------- the C++ facade
@pytest.fixture(scope="module")
def cpp_test(tmp_path_factory):
    import __graft_entry__ as ge

    ge.build_examples()  # makes sure libtitgpu.so and the headers are current
    exe = tmp_path_factory.mktemp("ttdb_cpp") / "test_data"
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "test_data.cpp"), "-o", str(exe),
                           f"-L{ROOT}/titsolver_b200", "-ltitgpu", f"-Wl,-rpath,{ROOT}/titsolver_b200", "-ldl"])
    return exe


def test_cpp_storage_known_answers_and_cross_reading(cpp_test, tmp_path):
    """The restated storage.test.cpp / type.test.cpp cases in C++; C++ reads a database
    written by Python, Python reads the one ParticleArray::write produced."""
    py = tmp_path / "python.ttdb"
    with ttdb.Storage(str(py)) as s:
        s.set_max_series(1)
        series = s.create_series("from python")
        n = 5
        r = np.arange(3.0 * n).reshape(n, 3)
        fields = {"r": r, "L": np.arange(9.0 * n).reshape(n, 3, 3), "rho": np.full(n, 1000.0, dtype=np.float32), "parinfo": np.arange(n, dtype=np.uint64)}
        series.write_particles(0.0, fields, names=list(fields))
        series.write_particles(1.5, fields, names=list(fields))
    out = subprocess.run([str(cpp_test), str(tmp_path), str(py)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok "), (out.stdout, out.stderr)
    with ttdb.Storage(str(tmp_path / "particles_cpp.ttdb"), read_only=True) as s:
        assert s.max_series == 1 and s.num_series == 1
        series = s.last_series()
        assert [f.time for f in series.frames()] == [0.0, 0.5]
        first, last = (f.read() for f in series.frames())
    assert list(first) == list(ttdb.PARTICLE_FIELDS)
    i = np.arange(7.0)
    assert np.array_equal(first["r"], np.stack([0.25 * i, 1.0 - 0.125 * i], axis=1))
    assert np.array_equal(first["v"], np.stack([i, -i], axis=1))
    assert np.array_equal(first["rho"], 1000.0 + i) and last["rho"][0] == 999.0 and np.array_equal(last["rho"][1:], first["rho"][1:])
    assert np.array_equal(first["L"], np.stack([np.stack([np.ones(7), i], axis=1), np.stack([-i, np.full(7, 2.0)], axis=1)], axis=1))
    assert first["grad_v"].shape == (7, 2, 2) and not first["grad_v"].any()


def test_cpp_reads_reference_blobs(cpp_test, tmp_path):
    """The C++ streaming decompressor on the reference's own frames: re-encode through
    C++ is not needed — a tiny driver reads the golden database and prints SHA-1-able bytes."""
    want = _storage_from_golden_blobs(tmp_path / "g.ttdb")
    src = tmp_path / "dump.cpp"
    src.write_text('''#include <cstdio>
#include "tit_b200/data.hpp"
int main(int, char** argv) {
  const tit::data::Storage s{argv[1], true};
  for (const auto& a : s.last_series().last_frame().arrays()) {
    const auto bytes = a.read();
    std::FILE* f = std::fopen((std::string{argv[2]} + "/" + a.name() + ".bin").c_str(), "wb");
    std::fwrite(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    std::printf("%s %s %zu\\n", a.name().c_str(), a.type().name().c_str(), a.size());
  }
}
''')
    exe = tmp_path / "dump"
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", str(src), "-o", str(exe), "-ldl"])
    out = subprocess.run([str(exe), str(tmp_path / "g.ttdb"), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert sorted(out.stdout.split("\n")[:2]) == sorted(f"{n} float64_t {a['size']}" for n, a in want.items())
    for name, a in want.items():
        assert hashlib.sha1((tmp_path / f"{name}.bin").read_bytes()).hexdigest() == a["sha1"]


# ---------------------------------------------------------------- GPU: frames of a real run
@pytest.mark.gpu
def test_solver_frames_roundtrip(tmp_path):
    """`particles.write(time, series)` of the reference's loop (wcsph.cpp:160, 185-188)
    from the Python binding: every field of a stepped solver survives the database."""
    import titsolver_b200 as tb

    case = tb.cases.dam_break_2d(20)
    gpu = tb.Solver(2)
    tb.load_case(gpu, case)
    gpu.initialize()
    with ttdb.Storage(str(tmp_path / "run.ttdb")) as s:
        s.set_max_series(1)
        series = s.create_series()
        ttdb.write_solver_frame(series, 0.0, gpu)
        gpu.step(3)
        ttdb.write_solver_frame(series, 1.0, gpu)
        frame = series.last_frame()
        assert [a.name for a in frame.arrays()] == list(ttdb.PARTICLE_FIELDS)
        got = frame.read()
        for f in ttdb.PARTICLE_FIELDS:
            assert np.array_equal(got[f], gpu.download(f)), f
        assert got["L"].shape == (case.n, 2, 2) and got["r"].shape == (case.n, 2)
        assert not np.array_equal(series.frame(0).read()["r"], got["r"])


@pytest.mark.gpu
def test_facade_driver_writes_ttdb(tmp_path):
    """examples/dam_break_2d.cpp with the reference's storage calls: the last frame of the
    database equals the raw dump of the final state."""
    import __graft_entry__ as ge

    exe = ge.build_examples()
    dump, db = tmp_path / "dump.bin", tmp_path / "particles.ttdb"
    r = subprocess.run([exe, "20", "5", str(dump), str(db)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with ttdb.Storage(str(db), read_only=True) as s:
        series = s.last_series()
        assert series.num_frames == 2 and series.frame(0).time == 0.0 and series.frame(1).time > 0.0
        first, last = series.frame(0).read(), series.frame(1).read()
    n = last["r"].shape[0]
    raw = np.fromfile(dump)
    assert np.array_equal(last["r"].ravel(), raw[: 2 * n]) and np.array_equal(last["v"].ravel(), raw[2 * n: 4 * n]) and np.array_equal(last["rho"], raw[4 * n:])
    assert list(last) == list(ttdb.PARTICLE_FIELDS) and last["L"].shape == (n, 2, 2)
    assert not first["v"].any() and last["v"].any()  # the initial frame is the state before the first step
    assert (last["gamma"] > 0).all() and np.isfinite(last["dv_dt"]).all() and last["dv_dt"].any()


# ---------------------------------------------------------------- the 3-D driver
def _driver_3d():
    import __graft_entry__ as ge

    return os.path.join(os.path.dirname(ge.build_examples()), "dam_break_3d")


def test_cpp_3d_driver_setup_matches_python_case(tmp_path):
    """examples/dam_break_3d.cpp builds the same tank as titsolver_b200.cases.dam_break_3d:
    wall and containment surfaces bit-identical, fluid within the jitter of the lattice."""
    from titsolver_b200 import cases

    db = tmp_path / "setup.ttdb"
    out = subprocess.run([_driver_3d(), "-5", "0", str(db), "0.93", "0.1"], capture_output=True, text=True)
    assert out.returncode == 0 and "200 fluid + 1890 fixed" in out.stdout, (out.stdout, out.stderr)
    with ttdb.Storage(str(db), read_only=True) as s:
        d = s.last_series().last_frame().read()
    c = cases.dam_break_3d(5, wall_ratio=0.93)
    assert np.array_equal(d["verts"], c.verts) and np.array_equal(d["faces"], c.faces)
    assert np.array_equal(d["containment_verts"], c.cverts) and np.array_equal(d["containment_faces"], c.cfaces)
    nf = c.n_fluid
    assert np.array_equal(d["r"][nf:], c.verts) and np.array_equal(d["m"], c.m) and np.array_equal(d["rho"][nf:], c.rho[nf:])
    off = np.abs(d["r"][:nf] - c.r[:nf]) / c.dr
    assert 0.05 < off.max() <= 0.1 and off.min() < 0.01
    assert np.allclose(d["rho"][:nf], c.rho0 + c.rho0 * c.g * (c.H - d["r"][:nf, 1]) / c.cs0**2, rtol=1e-15)
    # without jitter the lattice itself
    out = subprocess.run([_driver_3d(), "-4", "0", str(tmp_path / "s4.ttdb")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    with ttdb.Storage(str(tmp_path / "s4.ttdb"), read_only=True) as s:
        d4 = s.last_series().last_frame().read()
    c4 = cases.dam_break_3d(4)
    assert np.array_equal(d4["r"], c4.r) and np.array_equal(d4["rho"], c4.rho) and np.array_equal(d4["faces"], c4.faces)


@pytest.mark.gpu
def test_cpp_3d_driver_matches_python_solver_and_oracle(oracle, tmp_path):
    """Three SSPRK3 steps of the 3-D dam break through the C++ facade against the
    Python binding (same library: same bits) and against the oracle, both fed the
    particles of the driver's initial frame."""
    import dataclasses

    import titsolver_b200 as tb

    db = tmp_path / "run.ttdb"
    out = subprocess.run([_driver_3d(), "5", "3", str(db), "0.93", "0.1"], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    with ttdb.Storage(str(db), read_only=True) as s:
        series = s.last_series()
        assert series.num_frames == 2
        first, last = series.frame(0).read(), series.frame(1).read()
    case = dataclasses.replace(tb.cases.dam_break_3d(5, wall_ratio=0.93), r=first["r"], rho=first["rho"], m=np.full_like(first["m"], 1000.0 * (0.6 / 5) ** 3))
    nf = case.n_fluid
    assert np.array_equal(first["r"][nf:], case.verts)
    # initialize() scales the wall masses by gamma (fluid_equations.hpp:88); the frame holds the scaled ones
    assert np.array_equal(first["m"][:nf], case.m[:nf]) and (first["m"][nf:] <= case.m[nf:]).all()
    gpu = tb.Solver(3)
    tb.load_case(gpu, case)
    gpu.initialize()
    gpu.step(3)
    cpu = oracle.OracleSolver(3)
    oracle.load_case(cpu, case)
    cpu.initialize()
    cpu.step(3)
    for f, tol_same, tol_oracle in (("r", 1e-14, 1e-10), ("v", 1e-12, 1e-7), ("rho", 1e-14, 1e-10)):
        same, ref = gpu.download(f), cpu.download(f)
        assert np.abs(last[f] - same).max() <= tol_same * np.abs(same).max(), f
        assert np.abs(last[f][:nf] - ref[:nf]).max() <= tol_oracle * np.abs(ref[:nf]).max(), f
    assert last["L"].shape == (case.n, 3, 3) and last["v"][:nf].any()


# ---------------------------------------------------------------- ParaView export
def _check_xdmf(xdmf_path, series):
    import xml.etree.ElementTree as ET

    root = ET.parse(xdmf_path).getroot()
    assert root.tag == "Xdmf" and root.get("Version") == "3.0"
    (domain,) = list(root)
    (collection,) = list(domain)
    assert (collection.get("Name"), collection.get("GridType"), collection.get("CollectionType")) == ("TimeSeries", "Collection", "Temporal")
    frames = series.frames()
    grids = list(collection)
    assert len(grids) == len(frames)
    heavy = open(os.path.join(os.path.dirname(xdmf_path), "particles.bin"), "rb").read()

    def load(item):
        assert item.get("Format") == "Binary" and item.get("Endian") == "Little" and item.text == "particles.bin"
        shape = tuple(int(d) for d in item.get("Dimensions").split())
        dt = np.dtype({"Float": "f", "Int": "i"}[item.get("NumberType")] + item.get("Precision"))
        return np.frombuffer(heavy, dtype=dt, count=int(np.prod(shape)), offset=int(item.get("Seek"))).reshape(shape)

    width = len(str(len(frames) - 1)) if len(frames) > 1 else 1
    for index, (grid, frame) in enumerate(zip(grids, frames)):
        assert grid.get("GridType") == "Uniform" and grid.get("Name") == f"frame-{index:0{width}d}"
        data = frame.read()
        kids = list(grid)
        assert [k.tag for k in kids[:3]] == ["Time", "Topology", "Geometry"]
        assert float(kids[0].get("Value")) == frame.time
        assert kids[1].get("TopologyType") == "Polyvertex" and int(kids[1].get("NumberOfElements")) == data["r"].shape[0]
        assert kids[2].get("GeometryType") == {2: "XY", 3: "XYZ"}[data["r"].shape[1]]
        assert np.array_equal(load(kids[2][0]), data["r"])
        exported = [a for a in frame.arrays() if data[a.name].ndim < 3]  # matrices are skipped (hdf5.cpp:178-181)
        attrs = kids[3:]
        assert [a.get("Name") for a in attrs] == [a.name for a in exported]
        for attr in attrs:
            want = data[attr.get("Name")]
            assert attr.get("Center") == "Node" and attr.get("AttributeType") == ("Scalar" if want.ndim == 1 else "Vector")
            got = load(attr[0])
            assert got.shape == want.shape and np.array_equal(got.view(want.dtype) if got.dtype != want.dtype else got, want)


@pytest.mark.parametrize("dim, n_frames", [(2, 1), (3, 3), (2, 11)])
def test_xdmf_export(tmp_path, dim, n_frames):
    from titsolver_b200 import xdmf

    rng = np.random.default_rng(5)
    s = ttdb.Storage()
    series = s.create_series()
    for k in range(n_frames):
        n = 20 + k
        fields = {"r": rng.normal(size=(n, dim)), "rho": rng.normal(size=n), "v": rng.normal(size=(n, dim)), "L": rng.normal(size=(n, dim, dim)),
                  "parinfo": np.arange(n, dtype=np.uint64)}
        series.write_particles(0.5 * k, fields, names=["rho", "r", "L", "v", "parinfo"])
    out = xdmf.export_xdmf(str(tmp_path), series)
    assert os.path.basename(out) == "particles.xdmf"
    _check_xdmf(out, series)
    with pytest.raises(FileNotFoundError, match="Directory does not exist"):
        xdmf.export_xdmf(str(tmp_path / "nope"), series)
    with pytest.raises(NotADirectoryError, match="not a directory"):
        xdmf.export_xdmf(out, series)
    empty = s.create_series().create_frame(0.0)
    empty.create_array("rho").write(np.zeros(3))
    with pytest.raises(KeyError, match="'r' not found"):
        xdmf.export_xdmf(str(tmp_path), s.last_series())


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="the reference tree is only present in the build container")
def test_xdmf_export_of_the_reference_fixture(tmp_path):
    from titsolver_b200 import xdmf

    with ttdb.Storage(REF_FIXTURE, read_only=True) as s:
        series = s.last_series()
        _check_xdmf(xdmf.export_xdmf(str(tmp_path), series), series)


# ---------------------------------------------------------------- python -m titsolver_b200
def test_cli_arguments_and_no_cpu_fallback():
    from titsolver_b200 import __main__ as cli

    a = cli.parse(["--dim", "3", "--n-col", "16", "--max-steps", "7", "--kernel", "cubic", "--integrator", "verlet"])
    assert (a.dim, a.n_col, a.max_steps, a.frame_every, a.end_time, a.out) == (3, 16, 7, 100, 10.0, "particles.ttdb")
    assert cli.KERNELS.index(a.kernel) == 0 and cli.INTEGRATORS.index(a.integrator) == 1
    assert cli.KERNELS.index("wendland6") == 4 and cli.INTEGRATORS.index("ssprk3") == 3  # TITGPU_* ids of include/titgpu.h
    for bad in (["--n-col", "1"], ["--xdmf", "d", "--out", "-"], ["--dim", "1"]):
        with pytest.raises(SystemExit):
            cli.parse(bad)
    import torch

    if not torch.cuda.is_available():  # the product path fails loudly without a GPU
        import titsolver_b200 as tb

        with pytest.raises(tb.TitGpuError):
            cli.main(["--n-col", "10", "--max-steps", "2", "--out", "-"])


@pytest.mark.gpu
def test_cli_run_matches_direct_solver(tmp_path):
    import titsolver_b200 as tb
    from titsolver_b200 import __main__ as cli

    db, out = tmp_path / "particles.ttdb", tmp_path / "paraview"
    out.mkdir()
    assert cli.main(["--n-col", "20", "--max-steps", "5", "--frame-every", "2", "--out", str(db), "--xdmf", str(out)]) == 0
    case = tb.cases.dam_break_2d(20)
    gpu = tb.Solver(2)
    tb.load_case(gpu, case)
    gpu.initialize()
    for _ in range(5):
        gpu.step(1)
    with ttdb.Storage(str(db), read_only=True) as s:
        series = s.last_series()
        assert series.num_frames == 4  # initial, steps 2 and 4, the last step
        last = series.last_frame().read()
        for f in ("r", "v", "rho", "dv_dt", "L", "gamma"):
            assert np.array_equal(last[f], gpu.download(f)), f
        _check_xdmf(str(out / "particles.xdmf"), series)


def test_cpp_xdmf_export(tmp_path):
    """include/tit_b200/xdmf.hpp writes the same document as titsolver_b200/xdmf.py."""
    rng = np.random.default_rng(11)
    db = tmp_path / "run.ttdb"
    with ttdb.Storage(str(db)) as s:
        series = s.create_series("a <run> & more")
        for k in range(12):
            n = 30 + k
            fields = {"rho": rng.normal(size=n), "r": rng.normal(size=(n, 3)), "L": rng.normal(size=(n, 3, 3)), "v": rng.normal(size=(n, 3)),
                      "parinfo": np.arange(n, dtype=np.uint64), "kind": np.arange(n, dtype=np.int8)}
            series.write_particles(0.1 * k * np.pi, fields, names=list(fields))
    src = tmp_path / "export.cpp"
    src.write_text('''#include <cstdio>
#include "tit_b200/xdmf.hpp"
int main(int, char** argv) {
  try {
    const tit::data::Storage storage{argv[1], /*read_only=*/true};
    const auto out = tit::data::export_xdmf(argv[2], storage.last_series());
    std::printf("%s\\n", out.c_str());
    try { tit::data::export_xdmf(out, storage.last_series()); return 2; } catch (const tit::Exception&) {}
    try { tit::data::export_xdmf(std::string{argv[2]} + "/nope", storage.last_series()); return 3; } catch (const tit::Exception&) {}
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "%s\\n", e.what()); return 1; }
}
''')
    exe, out_dir = tmp_path / "export", tmp_path / "paraview"
    out_dir.mkdir()
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", str(src), "-o", str(exe), "-ldl"])
    run = subprocess.run([str(exe), str(db), str(out_dir)], capture_output=True, text=True)
    assert run.returncode == 0, (run.stdout, run.stderr)
    assert run.stdout.strip() == str(out_dir / "particles.xdmf")
    with ttdb.Storage(str(db), read_only=True) as s:
        _check_xdmf(str(out_dir / "particles.xdmf"), s.last_series())
