"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs. Bar (BASELINE.json north_star): neighbour sets bit-exact
after canonical sorting; density / acceleration / positions within 1e-10
relative (L-infinity scaled) for one evaluation or step, bounded drift over a
short horizon. drho_dt / dv_dt of fixed particles are never compared (the
reference accumulates garbage there)."""
import numpy as np
import pytest

import titsolver_b200 as tb
from titsolver_b200 import cases

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_err(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def make_pair(oracle, case, kernel_id=4, eos_id=0, integrator_id=3):
    g = tb.Solver(case.dim, kernel_id, eos_id, integrator_id)
    c = oracle.OracleSolver(case.dim, kernel_id, eos_id, integrator_id)
    tb.load_case(g, case)
    oracle.load_case(c, case)
    return g, c


def assert_csr_equal(a, b):
    (oa, ca), (ob, cb) = a, b
    assert np.array_equal(oa, ob), "row offsets differ"
    assert np.array_equal(ca, cb), "neighbour columns differ"


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("radius", [0.01, 0.1, 0.5, 1.0])
def test_neighbors_random_cloud(oracle, dim, radius):
    """geom/search.test.cpp:130-163: 200 random points, radii {0.01, 0.1, 0.5, 1}."""
    rng = np.random.default_rng(123)
    pts = rng.uniform(-1.0, 1.0, size=(200, dim))
    empty_v, empty_f = np.zeros((0, dim)), np.zeros((0, dim), np.uint64)
    res = []
    for S in (tb.Solver, oracle.OracleSolver):
        s = S(dim, 0)  # cubic spline: radius = 2h
        s.set_params(9.81, 1e-3, 10.0, 1000.0, 7.0, radius / 2.0)
        s.set_surface(empty_v, empty_f, empty_v, empty_f)
        s.set_particles(200, 0)
        s.upload("r", pts)
        res.append(s.neighbors())
    assert_csr_equal(res[0], res[1])
    # brute force, as the reference test does
    off, cols = res[0]
    for a in range(200):
        d2 = ((pts - pts[a]) ** 2).sum(1)
        assert set(cols[off[a]:off[a + 1]].tolist()) >= set(np.nonzero(d2 < (radius * (1 - 1e-12)) ** 2)[0].tolist())


@pytest.mark.parametrize("dim", [2, 3])
def test_neighbors_match_kd_tree_index(oracle, dim):
    """`geom::KDTreeSearch` is the reference's alternative index with the
    grid index's contract (search.test.cpp:122-125). On the GPU either option is served
    by the spatial hash; its rows equal the K-d tree's (oracle restatement) bit for bit."""
    rng = np.random.default_rng(321)
    pts = rng.uniform(0.0, 1.0, size=(3000, dim))
    empty_v, empty_f = np.zeros((0, dim)), np.zeros((0, dim), np.uint64)
    for radius in (0.02, 0.15):
        s = tb.Solver(dim, 0)
        s.set_params(9.81, 1e-3, 10.0, 1000.0, 7.0, radius / 2.0)
        s.set_surface(empty_v, empty_f, empty_v, empty_f)
        s.set_particles(len(pts), 0)
        s.upload("r", pts)
        assert_csr_equal(s.neighbors(), oracle.kdtree_neighbors(pts, radius))


@pytest.mark.parametrize("n_col", [20, 80])
def test_neighbors_dam_break_lattice(oracle, n_col):
    """The lattice puts neighbours at exactly 4 dr = 2h with an inclusive test."""
    case = cases.dam_break_2d(n_col)
    g, c = make_pair(oracle, case)
    assert_csr_equal(g.neighbors(), c.neighbors())
    off, _ = g.neighbors()
    cnt = np.diff(off.astype(np.int64))
    assert cnt[: case.n_fluid].max() == 49  # 49 incl. self on the 2-D lattice


def case_3d(n_col=6):
    """Generic positions: the 3-D wall integrals of the reference are evaluated at
    rounding-noise level when a face touches the support sphere exactly (see
    cases.dam_break_3d), so parity is checked off the lattice."""
    return cases.dam_break_3d(n_col, wall_ratio=0.93, jitter=0.1)


def test_neighbors_3d_lattice(oracle):
    case = cases.dam_break_3d(6)
    g, c = make_pair(oracle, case)
    assert_csr_equal(g.neighbors(), c.neighbors())


def test_initialize_gamma(oracle):
    case = cases.dam_break_2d(40)
    g, c = make_pair(oracle, case)
    g.initialize()
    c.initialize()
    for f in ("gamma", "grad_gamma", "m"):
        assert rel_err(g.download(f), c.download(f)) <= TOL, f
    gam = g.download("gamma")
    assert abs(gam[: case.n_fluid].max() - 1.0) < 1e-12


def check_rhs(oracle, case, **kw):
    g, c = make_pair(oracle, case, **kw)
    g.initialize()
    c.initialize()
    g.rhs_only()
    c.rhs_only()
    nf = case.n_fluid
    for f in ("gamma", "grad_gamma", "rho", "p", "cs"):
        assert rel_err(g.download(f), c.download(f)) <= TOL, f
    for f in ("drho_dt", "dv_dt"):
        assert rel_err(g.download(f)[:nf], c.download(f)[:nf]) <= TOL, f
    return g, c


def test_rhs_parity_2d(oracle):
    check_rhs(oracle, cases.dam_break_2d(40))


@pytest.mark.parametrize("kernel_id", [0, 1, 2, 3, 5])
def test_rhs_parity_2d_all_kernels(oracle, kernel_id):
    check_rhs(oracle, cases.dam_break_2d(16), kernel_id=kernel_id)


def test_rhs_parity_2d_linear_eos(oracle):
    check_rhs(oracle, cases.dam_break_2d(16), eos_id=1)


def test_rhs_parity_2d_moving(oracle):
    """Non-trivial velocities and a perturbed lattice."""
    case = cases.dam_break_2d(24)
    rng = np.random.default_rng(7)
    nf = case.n_fluid
    case.r[:nf] += rng.uniform(-0.2, 0.2, size=(nf, 2)) * case.dr
    v = np.zeros_like(case.r)
    v[:nf] = rng.normal(size=(nf, 2))
    g, c = make_pair(oracle, case)
    g.upload("v", v)
    c.upload("v", v)
    g.initialize(); c.initialize()
    g.rhs_only(); c.rhs_only()
    for f in ("drho_dt", "dv_dt"):
        assert rel_err(g.download(f)[:nf], c.download(f)[:nf]) <= TOL, f


STEP_FIELDS = ("r", "v", "rho")
POST_FIELDS = ("N", "L", "grad_v", "grad_rho", "dr", "phi", "rho_raw", "gamma", "grad_gamma", "drho_dt", "dv_dt", "p", "cs")


def check_step(oracle, case, nsteps=1, tol=TOL, **kw):
    g, c = make_pair(oracle, case, **kw)
    g.initialize(); c.initialize()
    dt_g = g.step(nsteps)
    dt_c = c.step(nsteps)
    assert abs(dt_g - dt_c) <= 1e-12 * dt_c
    nf = case.n_fluid
    for f in STEP_FIELDS:
        assert rel_err(g.download(f), c.download(f)) <= tol, f
    return g, c


def test_one_step_parity_2d(oracle):
    case = cases.dam_break_2d(40)
    g, c = check_step(oracle, case)
    nf = case.n_fluid
    for f in POST_FIELDS:
        a, b = g.download(f), c.download(f)
        if f in ("drho_dt", "dv_dt"):
            a, b = a[:nf], b[:nf]
        assert rel_err(a, b) <= 1e-9, f


@pytest.mark.parametrize("integrator_id", [0, 1, 2])
def test_one_step_other_integrators(oracle, integrator_id):
    check_step(oracle, cases.dam_break_2d(16), integrator_id=integrator_id)


def test_drift_20_steps_2d(oracle):
    """Bounded drift over a short horizon (chaotic amplification of rounding)."""
    check_step(oracle, cases.dam_break_2d(20), nsteps=20, tol=1e-7)


def test_rhs_parity_3d(oracle):
    check_rhs(oracle, case_3d())


def test_one_step_parity_3d(oracle):
    check_step(oracle, case_3d())


def test_rhs_parity_3d_shared_edges_second_mesh(oracle):
    """Every interior edge of the structured wall mesh has a coplanar twin, so the 3-D
    wall pipeline evaluates it once (flux) or not at all (antigradient rim sum); a second
    mesh / lattice ratio besides case_3d(). (Ratio 1 is not a parity case: wall particles
    then see faces that touch their support sphere exactly, see cases.dam_break_3d; the same
    holds whenever the support radius is a multiple of the wall spacing, e.g. ratio 0.81 here.)"""
    check_rhs(oracle, cases.dam_break_3d(6, wall_ratio=0.77, jitter=0.13))


def test_rhs_parity_3d_fine_wall_mesh_overflows_the_search_stage(oracle):
    """A wall mesh 2.3x finer than the particle spacing puts > 224 faces within reach of
    a near-wall particle: the search stage hands those particles to the generic kernel."""
    case = cases.dam_break_3d(5, wall_ratio=0.43, jitter=0.1)
    g, c = make_pair(oracle, case)
    g.initialize()
    c.initialize()
    g.profile(True)
    g.profile_reset()
    g.rhs_only()
    c.rhs_only()
    names = " ".join(g.profile_read())
    assert "k_wsearch" in names and "k_wall" in names, names
    nf = case.n_fluid
    for f in ("gamma", "grad_gamma", "rho"):
        assert rel_err(g.download(f), c.download(f)) <= TOL, f
    for f in ("drho_dt", "dv_dt"):
        assert rel_err(g.download(f)[:nf], c.download(f)[:nf]) <= TOL, f


def test_rhs_and_step_parity_3d_slanted_refined_walls(oracle):
    """A tetrahedral tank tessellated by red refinement: slanted faces, irregular
    triangles, coplanar twins whose normals agree only to rounding."""
    case = cases.tetra_tank_3d(14, fill=0.7)
    assert case.n_fluid > 50 and len(case.faces) > 500
    check_rhs(oracle, case)
    check_step(oracle, case)


@pytest.mark.parametrize("kernel_id", [0, 2])
def test_rhs_parity_3d_other_kernels(oracle, kernel_id):
    check_rhs(oracle, case_3d(5), kernel_id=kernel_id)


def test_rhs_parity_3d_linear_eos(oracle):
    check_rhs(oracle, case_3d(5), eos_id=1)


@pytest.mark.parametrize("dim", [2, 3])
def test_one_step_parity_general_tait_exponent(oracle, dim):
    """xi != 7: the pair loop gathers the neighbours' EOS record instead of recomputing it."""
    case = cases.dam_break_2d(16) if dim == 2 else case_3d(5)
    case.xi = 5.0
    check_step(oracle, case)


def test_output_levels_do_not_change_the_state(oracle):
    """titgpu_set_outputs: publishing derived fields is optional, the state evolution is not."""
    case = cases.dam_break_2d(16)
    ref = None
    for level in (2, 1, 0):
        g = tb.Solver(2)
        tb.load_case(g, case)
        g.set_outputs(level)
        g.initialize()
        g.step(3)
        st = [g.download(f) for f in STEP_FIELDS]
        if ref is None:
            ref = st
            N2 = g.download("N")
            assert np.abs(N2[case.n_fluid:]).max() > 0  # wall-particle sums are published at level 2
        else:
            for a, b in zip(st, ref):
                assert np.array_equal(a, b)
        if level == 1:
            N1 = g.download("N")
            assert np.array_equal(N1[: case.n_fluid], N2[: case.n_fluid])
            assert np.abs(N1[case.n_fluid:]).max() == 0
        if level == 0:
            assert np.abs(g.download("N")).max() == 0
    with pytest.raises(tb.TitGpuError):
        g.set_outputs(3)


def _run(case, lists, v=None, nsteps=3, dim=2):
    g = tb.Solver(dim)
    tb.load_case(g, case)
    g.set_lists(lists)
    if v is not None:
        g.upload("v", v)
    g.initialize()
    g.step(nsteps)
    return g, [g.download(f) for f in STEP_FIELDS]


@pytest.mark.parametrize("dim", [2, 3])
def test_candidate_lists_match_search_at_every_prepare(dim):
    """titgpu_set_lists: one skin-enlarged search per step vs the reference's four;
    the neighbour sets are the same, only the order of the sums differs."""
    case = cases.dam_break_2d(24) if dim == 2 else case_3d()
    g1, a = _run(case, True, dim=dim)
    g0, b = _run(case, False, dim=dim)
    assert g1.list_redos == 0
    for x, y in zip(a, b):
        assert rel_err(x, y) <= 1e-11


def test_candidate_lists_fall_back_when_the_skin_is_exceeded():
    """Near-sonic velocities move particles by more than skin / 2 in one step: the
    call is repeated with a search at every prepare and gives that result (up to the
    order of the sums: the hash cells are a skin wider when lists are enabled)."""
    case = cases.dam_break_2d(16)
    rng = np.random.default_rng(5)
    v = np.zeros_like(case.r)
    v[: case.n_fluid] = rng.normal(size=(case.n_fluid, 2)) * case.cs0
    g1, a = _run(case, True, v, nsteps=2)
    g0, b = _run(case, False, v, nsteps=2)
    assert g1.list_redos == 1 and g0.list_redos == 0
    for x, y in zip(a, b):
        assert rel_err(x, y) <= 1e-9


def test_upload_of_r_keeps_or_refreshes_the_wall_cache(oracle):
    """gamma / grad gamma of the wall particles are cached. An upload of `r` that leaves
    them in place (every upload of the reference's time loop) keeps the cache; moving one
    of them rebuilds it - and the results follow the oracle, which has no cache."""
    case = cases.dam_break_2d(16)
    g, c = make_pair(oracle, case)
    g.initialize(); c.initialize()
    g.step(1)
    state = {f: g.download(f) for f in ("r", "v", "rho")}
    for f, a in state.items():
        c.upload(f, a)

    def launches_of_rhs():
        n0 = g.launch_count
        g.rhs_only()
        return g.launch_count - n0

    def compare():
        c.rhs_only()
        nf = case.n_fluid
        for f in ("gamma", "grad_gamma", "rho"):
            assert rel_err(g.download(f), c.download(f)) <= TOL, f
        for f in ("drho_dt", "dv_dt"):
            assert rel_err(g.download(f)[:nf], c.download(f)[:nf]) <= TOL, f

    launches_of_rhs()
    base = launches_of_rhs()
    g.upload("r", state["r"])
    assert launches_of_rhs() == base  # cache kept
    compare()
    moved = state["r"].copy()
    moved[case.n_fluid + 7] += 0.3 * case.dr * np.array([1.0, 1.0])
    g.upload("r", moved)
    c.upload("r", moved)
    assert launches_of_rhs() == base + 1  # cache rebuilt (2-D: one more wall-kernel launch)
    compare()


def test_strided_upload_download(oracle):
    """The reference pads Vec<double,3> to 32 bytes."""
    case = cases.dam_break_3d(4)
    g = tb.Solver(3)
    tb.load_case(g, case)
    padded = np.zeros((case.n, 4))
    padded[:, :3] = case.r
    g.upload("r", padded, stride_bytes=32)
    out = np.full((case.n, 4), -1.0)
    g.download_raw("r", out.ctypes.data, 32)
    assert np.array_equal(out[:, :3], case.r)
    assert (out[:, 3] == -1.0).all()


def test_errors_are_reported():
    with pytest.raises(tb.TitGpuError):
        tb.Solver(4)
    s = tb.Solver(2)
    with pytest.raises(tb.TitGpuError):
        s.step(1)  # nothing set up yet


def test_cuda_graph_replay_is_bit_identical_2d():
    """titgpu_set_graphs: whole 2-D steps recorded once per ping-pong state (period three) and
    replayed; nothing about the results may change, whatever the mix of step() calls, output
    levels and uploads in between."""
    case = cases.dam_break_2d(24)
    res = []
    for graphs in (True, False):
        g = tb.Solver(2)
        g.set_graphs(graphs)
        tb.load_case(g, case)
        g.initialize()
        g.step(1)
        g.step(7)
        g.set_outputs(1)
        g.step(2)
        r = g.download("r")
        g.upload("r", r)  # an upload that moves nothing keeps the graphs valid
        g.set_outputs(2)
        g.step(4)
        res.append({f: g.download(f) for f in STEP_FIELDS + ("N", "phi", "dv_dt", "p")})
        res[-1]["launches"] = g.launch_count
        assert (g.graph_replays > 0) == graphs
    for f in STEP_FIELDS + ("N", "phi", "dv_dt", "p"):
        assert np.array_equal(res[0][f], res[1][f]), f
    assert res[0]["launches"] == res[1]["launches"]  # replayed launches are counted too


@pytest.mark.parametrize("dim", [2, 3])
def test_published_fields_of_dry_wall_particles_are_kept_not_recomputed(oracle, dim):
    """Output level 2 publishes N, L, grad_v, grad_rho, dr, phi of the wall particles although the
    step never reads them. Far from any fluid they are static, and the output pass skips them after
    the first publish (k_deep_dry): same values as the oracle at every publishing step, fewer
    launches' worth of work - and an upload that could change them forces a recomputation."""
    case = cases.dam_break_2d(24) if dim == 2 else cases.dam_break_3d(8, wall_ratio=0.93, jitter=0.1)
    g, c = make_pair(oracle, case)
    g.set_graphs(False)
    g.initialize(); c.initialize()
    nf = case.n_fluid
    for k in range(3):
        g.step(2); c.step(2)
        for f in ("N", "L", "grad_v", "grad_rho", "dr", "phi", "gamma", "grad_gamma", "rho", "p"):
            assert rel_err(g.download(f)[nf:], c.download(f)[nf:]) <= 1e-9, (k, f)
    # the same with the cache off: bit-identical published fields
    import os

    os.environ["TITGPU_DRY_CACHE"] = "0"
    try:
        g2 = tb.Solver(dim)
        g2.set_graphs(False)
        tb.load_case(g2, case)
        g2.initialize()
        for k in range(3):
            g2.step(2)
    finally:
        del os.environ["TITGPU_DRY_CACHE"]
    for f in ("N", "L", "grad_v", "grad_rho", "dr", "phi", "r", "v", "rho"):
        assert np.array_equal(g.download(f), g2.download(f)), f
    # a new wall density changes what the neighbours' sums are made of: recomputed, still equal to the oracle
    rho = g.download("rho")
    rho[nf:] *= 1.0 + 1e-3
    g.upload("rho", rho); c.upload("rho", rho)
    g.step(1); c.step(1)
    for f in ("N", "grad_rho", "dr"):
        assert rel_err(g.download(f)[nf:], c.download(f)[nf:]) <= 1e-9, f


@pytest.mark.gpu
@pytest.mark.parametrize("dim,kernel_id,eos_id", [(2, 4, 0), (3, 4, 0), (3, 2, 1), (2, 1, 0)])
def test_grouped_candidate_sweep_is_bit_identical(dim, kernel_id, eos_id):
    """titgpu_set_group_sweep: k_rhs_grp / k_shift_grp (one candidate sweep per 4 consecutive
    particles, one compacted hit list per particle) keep the order of every sum of the default
    traversal, so forcing them on a small case must reproduce it BIT for bit - state and all
    published fields (the output pass includes the wall particles), over steps that move
    particles between cells. A spray of loose particles exercises short, empty and ragged runs."""
    case = cases.dam_break_2d(30) if dim == 2 else cases.dam_break_3d(9, wall_ratio=0.93, jitter=0.1)
    rng = np.random.default_rng(5)
    v = np.zeros((case.n, dim))
    v[: case.n_fluid] = 0.8 * rng.standard_normal((case.n_fluid, dim))
    fields = ("r", "v", "rho", "drho_dt", "dv_dt", "p", "cs", "gamma", "grad_gamma", "N", "L", "grad_v", "grad_rho", "dr", "phi")
    res = []
    for mode in (1, 0):
        g = tb.Solver(dim, kernel_id, eos_id)
        g.set_graphs(False)
        g.set_group_sweep(mode)
        tb.load_case(g, case)
        g.upload("v", v)
        g.initialize()
        g.rhs_only()
        out = {"rhs_" + f: g.download(f) for f in ("drho_dt", "dv_dt")}
        g.step(1)
        g.step(6)
        out.update({f: g.download(f) for f in fields})
        res.append(out)
    for f in res[0]:
        assert np.array_equal(res[0][f], res[1][f], equal_nan=True), f


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_grouped_sweep_falls_back_when_a_hit_list_overflows(oracle, dim):
    """A smoothing length far above the lattice spacing gives every particle more neighbours than a
    member's list of the grouped sweep holds (384; here ~500 in 2-D, ~1100 in 3-D): the group must
    leave the sweep half-way and take the gather traversal, which in turn drains its own list
    (512 entries) in the middle of the sweep. Same results bit for bit as with the grouped sweep off,
    and the oracle agrees on the right-hand sides."""
    import dataclasses

    base = cases.dam_break_2d(20) if dim == 2 else cases.dam_break_3d(8, wall_ratio=0.93, jitter=0.1)
    case = dataclasses.replace(base, h=base.h * (3.2 if dim == 2 else 1.6))
    res = []
    for mode in (1, 0):
        g = tb.Solver(dim)
        g.set_graphs(False)
        g.set_group_sweep(mode)
        tb.load_case(g, case)
        g.initialize()
        g.rhs_only()
        out = {"rhs_" + f: g.download(f) for f in ("drho_dt", "dv_dt")}
        g.step(2)
        out.update({f: g.download(f) for f in STEP_FIELDS + ("N", "phi", "grad_v")})
        res.append(out)
    for f in res[0]:
        assert np.array_equal(res[0][f], res[1][f], equal_nan=True), f
    off, _ = g.neighbors()
    assert np.diff(off)[: case.n_fluid].max() > 384
    c = oracle.OracleSolver(dim, 4, 0, 3)
    oracle.load_case(c, case)
    c.initialize()
    c.rhs_only()
    nf = case.n_fluid
    for f in ("drho_dt", "dv_dt"):
        assert rel_err(res[0]["rhs_" + f][:nf], c.download(f)[:nf]) < 1e-9, f
