"""Pins of the oracle's PHYSICS layer (FluidEquations), which no reference test covers.

(1) An independent transcription of the reference's pair lambdas in their original
    SYMMETRIC SCATTER form — one visit per unordered pair (a, b), b < a, updating both
    particles, exactly as `par::block_for_each(mesh.block_pairs(...))` does
    (/root/reference/source/tit/sph/fluid_equations.hpp:249-259 continuity, :293-304
    momentum, :351-365 shifting sums, :366-377 renormalisation, :396-416 visibility,
    :419-452 splash / near-surface, :455-470 shifts, :489-511 free-surface correction)
    — written in numpy from the reference source, sharing no code with oracle/ (own
    kernel formulas, own neighbour search, own LU). The oracle evaluates the same sums in
    GATHER form (every ordered pair from a's side); the two must agree to rounding.
(2) The sanity values DESIGN.md section 5 claims for the wall integrals: gamma = 1 in the
    interior, 1/2 on a flat wall, 1/4 in a 2-D corner (kernel.test.cpp:376-393 checks the
    same closed forms for single faces), and a column in hydrostatic equilibrium feels no
    net force.
"""
import math

import numpy as np
import pytest

import oracle_lib
from titsolver_b200 import cases

TINY = np.finfo(np.float64).eps ** (1.0 / 3.0)  # core/math.hpp:145-147
PHI_MIN, PHI_MAX = np.finfo(np.float64).tiny, 1.0


# ---- Wendland C4 ("SixthOrderWendland"), kernel.gen.cpp:520-554 / SURVEY App. B ----
def w_unit(q):
    return np.where(q < 2.0, (1.0 + 3.0 * q + 35.0 / 12.0 * q * q) * (1.0 - q / 2.0) ** 6, 0.0)


def dw_unit(q):
    return np.where(q < 2.0, 7.0 * q * (q - 2.0) ** 5 * (5.0 * q + 2.0) / 96.0, 0.0)


def omega(dim):
    return 9.0 / (4.0 * math.pi) if dim == 2 else 495.0 / (256.0 * math.pi)


def W(x, h):  # kernel.hpp:134-141
    dim = x.shape[-1]
    q = np.sqrt((x * x).sum(-1)) / h
    return omega(dim) * h ** (-dim) * w_unit(q)


def gradW(x, h):  # kernel.hpp:154-163; normalize() -> 0 below tiny
    dim = x.shape[-1]
    d = np.sqrt((x * x).sum(-1))
    xhat = np.where((d >= TINY)[..., None], x / np.maximum(d, 1e-300)[..., None], 0.0)
    return (omega(dim) * h ** (-dim) * dw_unit(d / h) / h)[..., None] * xhat


def eos(rho, rho0, cs0, xi=7.0):  # equation_of_state.hpp:19-72
    return rho0 * cs0**2 / xi * ((rho / rho0) ** xi - 1.0), cs0 * (rho / rho0) ** ((xi - 1.0) / 2.0)


def pairs_within(r, radius):
    """Unordered pairs (a, b), b < a, |r_a - r_b|^2 <= radius^2 (brute force)."""
    d2 = ((r[:, None, :] - r[None, :, :]) ** 2).sum(-1)
    a, b = np.nonzero(np.tril(d2 <= radius * radius, k=-1))
    return a, b


def lu_inverse_nopivot(A):
    """core/_mat/fact.hpp:84-108: LU without pivoting, fails on a tiny pivot."""
    n = A.shape[0]
    LU = np.zeros_like(A)
    for i in range(n):
        for j in range(i):
            s = A[i, j] - sum(LU[i, k] * LU[k, j] for k in range(j))
            LU[i, j] = s / LU[j, j]
        for j in range(i, n):
            LU[i, j] = A[i, j] - sum(LU[i, k] * LU[k, j] for k in range(i))
        if abs(LU[i, i]) <= TINY:
            return None
    Lm, U = np.tril(LU, -1) + np.eye(n), np.triu(LU)
    return np.linalg.solve(U, np.linalg.solve(Lm, np.eye(n)))


def blob(dim, n_side, seed, velocity=1.0, h=0.05):
    """A free blob of jittered particles: no walls, gamma = 1 everywhere."""
    rng = np.random.default_rng(seed)
    dr = h / 2.0
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * dim, indexing="ij"), -1).reshape(-1, dim).astype(float)
    r = (g + 0.5 + rng.uniform(-0.2, 0.2, g.shape)) * dr
    v = rng.normal(size=r.shape) * velocity
    rho = 1000.0 * (1.0 + 0.01 * rng.normal(size=len(r)))
    m = np.full(len(r), 1000.0 * dr**dim) * (1.0 + 0.05 * rng.uniform(-1, 1, len(r)))
    return r, v, rho, m, h


def oracle_blob(dim, r, v, rho, m, h, g=9.81, mu=1e-3, cs0=30.0, rho0=1000.0):
    s = oracle_lib.OracleSolver(dim)
    s.set_params(g, mu, cs0, rho0, 7.0, h)
    e = np.zeros((0, dim))
    big = 1e3  # containment: a huge box, so that gamma = 1 without any boundary face
    if dim == 2:
        cv = np.array([[-big, -big], [big, -big], [big, big], [-big, big]])
        cf = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], np.uint64)
    else:
        cv, cf = cases._box_wall_mesh((2 * big,) * 3, (1, 1, 1), inward=False)
        cv = cv - big
    s.set_surface(e, e.astype(np.uint64), cv, cf)
    s.set_particles(len(r), 0)
    for f, a in (("r", r), ("v", v), ("rho", rho), ("m", m)):
        s.upload(f, a)
    return s


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("dim,n_side", [(2, 24), (3, 9)])
def test_scatter_form_rhs_equals_gather_form_oracle(dim, n_side):
    r, v, rho, m, h = blob(dim, n_side, seed=11 + dim)
    g, mu, cs0, rho0 = 9.81, 1e-3, 30.0, 1000.0
    orc = oracle_blob(dim, r, v, rho, m, h, g, mu, cs0, rho0)
    orc.rhs_only()
    assert np.all(orc.download("gamma") == 1.0)

    a, b = pairs_within(r, 2.0 * h)
    assert len(a) > 10 * len(r)
    p, cs = eos(rho, rho0, cs0)
    r_ab, v_ab, rho_ab = r[a] - r[b], v[a] - v[b], rho[a] - rho[b]
    gW = gradW(r_ab, h)
    # continuity, fluid_equations.hpp:249-259
    cs_ab = np.maximum(cs[a], cs[b])
    Psi = (cs_ab * rho_ab / np.sqrt((r_ab**2).sum(-1)))[:, None] * r_ab
    drho = np.zeros(len(r))
    np.add.at(drho, a, m[b] * ((v_ab + Psi / rho[b][:, None]) * gW).sum(-1))
    np.add.at(drho, b, -m[a] * ((-v_ab + Psi / rho[a][:, None]) * gW).sum(-1))
    # momentum, :293-304
    P_ab = p[a] / rho[a] ** 2 + p[b] / rho[b] ** 2
    Pi_ab = 2.0 * mu * (v_ab * r_ab).sum(-1) / (rho[a] * rho[b] * (r_ab**2).sum(-1))
    dv = np.zeros_like(r)
    dv[:, 1] = -g
    np.add.at(dv, a, (m[b] * (Pi_ab - P_ab))[:, None] * gW)
    np.add.at(dv, b, -(m[a] * (Pi_ab - P_ab))[:, None] * gW)

    assert rel(drho, orc.download("drho_dt")) <= 1e-13
    assert rel(dv, orc.download("dv_dt")) <= 1e-13
    assert rel(p, orc.download("p")) <= 1e-14 and rel(cs, orc.download("cs")) <= 1e-14


def scatter_post(r, v, rho, m, h):
    """apply_shifts + apply_free_surface_correction (gamma = 1, no faces), scatter form."""
    n, dim = r.shape
    radius = 2.0 * h
    a, b = pairs_within(r, radius)
    r_ab = r[a] - r[b]
    gW = gradW(r_ab, h)
    V = m / rho
    N = np.zeros((n, dim)); L = np.zeros((n, dim, dim)); gv = np.zeros((n, dim, dim)); gr = np.zeros((n, dim))
    outer = lambda x, y: x[:, :, None] * y[:, None, :]  # noqa: E731
    # :351-365
    np.add.at(N, a, V[b][:, None] * gW)
    np.add.at(N, b, -V[a][:, None] * gW)
    np.add.at(L, a, V[b][:, None, None] * outer(-r_ab, gW))
    np.add.at(L, b, -V[a][:, None, None] * outer(r_ab, gW))
    np.add.at(gv, a, V[b][:, None, None] * outer(v[b] - v[a], gW))
    np.add.at(gv, b, -V[a][:, None, None] * outer(v[a] - v[b], gW))
    np.add.at(gr, a, (V[b] * (rho[b] - rho[a]))[:, None] * gW)
    np.add.at(gr, b, -(V[a] * (rho[a] - rho[b]))[:, None] * gW)
    # :366-377
    dr = N.copy()
    lu_failed = 0
    for i in range(n):
        inv = lu_inverse_nopivot(L[i].T.copy())
        if inv is not None:
            L[i] = inv
            N[i] = L[i] @ N[i]
            gv[i] = gv[i] @ L[i].T
            gr[i] = L[i] @ gr[i]
        else:
            L[i] = np.eye(dim)
            lu_failed += 1
        nn = np.sqrt((N[i] ** 2).sum())
        N[i] = N[i] / nn if nn >= TINY else 0.0
    # :387-416 (all particles are fluid here)
    phi = np.full(n, PHI_MIN)
    cos2 = math.cos(math.pi / 4) ** 2
    r2 = (r_ab**2).sum(-1)
    n_a, n_b = (N[a] * r_ab).sum(-1), (N[b] * r_ab).sum(-1)
    phi[a[(n_a > 0) & (n_a**2 >= cos2 * r2)]] = PHI_MAX
    phi[b[(n_b < 0) & (n_b**2 >= cos2 * r2)]] = PHI_MAX
    # :419-426 (adjacency includes the particle itself)
    count = 1 + np.bincount(a, minlength=n) + np.bincount(b, minlength=n)
    phi[count <= (8 if dim == 2 else 26)] = PHI_MIN
    # :440-452
    on_fs = phi == PHI_MIN
    phi2 = phi.copy()
    d2_full = ((r[:, None, :] - r[None, :, :]) ** 2).sum(-1)
    for i in np.nonzero(phi == PHI_MAX)[0]:
        nb = np.nonzero((d2_full[i] <= radius * radius) & on_fs)[0]
        if len(nb):
            j = nb[np.argmin(d2_full[i, nb])]  # first minimum = lowest index on ties
            phi2[i] = phi[i] * abs(np.dot(N[j], r[i] - r[j])) / radius
    phi = phi2
    # :455-470
    r2_, v2, rho2 = r.copy(), v.copy(), rho.copy()
    far = phi == PHI_MAX
    dr[~far] = 0.0
    dr[far] *= -0.4 * 0.2 * h * h
    r2_ += dr
    v2[far] += np.einsum("nij,nj->ni", gv[far], dr[far])
    rho2[far] += (gr[far] * dr[far]).sum(-1)
    # :489-511 — neighbour sets are those of the pre-shift mesh, positions the shifted ones
    rho_raw = rho2.copy()
    rho3 = rho2.copy()
    K_fs = -math.log(0.05) / 0.01**2
    corrected = 0
    for i in np.nonzero(~far)[0]:
        nb = np.nonzero(d2_full[i] <= radius * radius)[0]
        Wv = W(r2_[i] - r2_[nb], h)
        alpha, rho_t = (m[nb] / rho_raw[nb] * Wv).sum(), (m[nb] * Wv).sum()
        ratio = min(1.0, alpha / 1.0)
        if ratio > 0.99:
            continue
        beta = math.exp(-K_fs * (ratio - 1.0) ** 2)
        corr = beta * 1.0 + (1.0 - beta) * alpha
        if abs(corr) > TINY:
            rho3[i] = rho_t / corr
            corrected += 1
    return dict(N=N, L=L, grad_v=gv, grad_rho=gr, phi=phi, dr=dr, r=r2_, v=v2, rho=rho3, rho_raw=rho_raw), dict(lu_failed=lu_failed, corrected=corrected)


@pytest.mark.parametrize("dim,n_side", [(2, 22), (3, 12)])
def test_scatter_form_post_integrate_equals_oracle(dim, n_side):
    r, v, rho, m, h = blob(dim, n_side, seed=5 + dim, velocity=0.3)
    orc = oracle_blob(dim, r, v, rho, m, h)
    orc.post_only()
    st = orc.stats()
    ref, cnt = scatter_post(r, v, rho, m, h)
    # every branch class is populated in this case: free surface, near surface, interior
    assert st["free_surface"] > 0 and st["near_surface"] > 0 and st["shifted"] > 0 and st["fs_corrected"] > 0
    assert st["lu_failed"] == cnt["lu_failed"] and st["fs_corrected"] == cnt["corrected"]
    phi_o = orc.download("phi")
    assert np.array_equal(phi_o == PHI_MIN, ref["phi"] == PHI_MIN) and np.array_equal(phi_o == PHI_MAX, ref["phi"] == PHI_MAX)
    for f, tol in (("N", 1e-10), ("L", 1e-10), ("grad_v", 1e-10), ("grad_rho", 1e-10), ("phi", 1e-10), ("dr", 1e-10), ("r", 1e-14), ("v", 1e-12), ("rho", 1e-12),
                   ("rho_raw", 1e-13)):
        assert rel(orc.download(f), ref[f]) <= tol, f


def test_lu_failure_branch_gives_identity():
    """An isolated pair: L is rank one, the LU of L^T meets a tiny pivot and the reference
    falls back to L = I (fluid_equations.hpp:374-376); two neighbours <= 8 -> splash rule."""
    h = 0.05
    r = np.array([[0.0, 0.0], [0.6 * h, 0.0]])
    v, rho, m = np.zeros_like(r), np.full(2, 1000.0), np.full(2, 1000.0 * (h / 2) ** 2)
    orc = oracle_blob(2, r, v, rho, m, h)
    orc.post_only()
    st = orc.stats()
    assert st["lu_failed"] == 2 and st["free_surface"] == 2
    assert np.array_equal(orc.download("L"), np.broadcast_to(np.eye(2), (2, 2, 2)))
    ref, cnt = scatter_post(r, v, rho, m, h)
    assert cnt["lu_failed"] == 2
    assert rel(orc.download("rho"), ref["rho"]) <= 1e-13


# ---- (2) wall integrals and equilibrium ------------------------------------------------
def test_gamma_interior_wall_corner_2d():
    case = cases.dam_break_2d(40)
    s = oracle_lib.OracleSolver(2)
    oracle_lib.load_case(s, case)
    s.initialize()
    gamma, gg = s.download("gamma"), s.download("grad_gamma")
    nf = case.n_fluid
    rf = case.r[:nf]
    radius = 2.0 * case.h
    PW, PH = case.meta["tank"]
    interior = (rf[:, 0] > radius) & (rf[:, 1] > radius)
    assert interior.sum() > 100 and np.all(gamma[:nf][interior] == 1.0) and np.all(gg[:nf][interior] == 0.0)
    # wall particles: the four tank corners see a quarter of the kernel support, wall
    # particles farther than a support radius from any corner see half of it
    rx = case.r[nf:]
    corner = ((np.abs(rx[:, 0]) < 1e-12) | (np.abs(rx[:, 0] - PW) < 1e-12)) & ((np.abs(rx[:, 1]) < 1e-12) | (np.abs(rx[:, 1] - PH) < 1e-12))
    assert corner.sum() == 4
    dcorner = np.minimum.reduce([np.hypot(rx[:, 0] - cx, rx[:, 1] - cy) for cx in (0.0, PW) for cy in (0.0, PH)])
    flat = dcorner > radius
    assert flat.sum() > 100
    assert np.abs(gamma[nf:][flat] - 0.5).max() <= 0.03  # evaluated h^2 = 0.06 h off the wall (fluid_equations.hpp:187): 1/2 +- 0.025
    assert np.abs(gamma[nf:][corner] - 0.25).max() <= 0.03
    # a fluid particle half a spacing... the first fluid row sits one spacing above the floor:
    # its gamma lies strictly between 1/2 and 1 and grad gamma points away from the wall
    row = (np.abs(rf[:, 1] - case.dr) < 1e-12) & (rf[:, 0] > radius) & (rf[:, 0] < 2 * case.H - radius)
    assert np.all((gamma[:nf][row] > 0.5) & (gamma[:nf][row] < 1.0))
    assert np.all(gg[:nf][row][:, 1] > 0) and np.abs(gg[:nf][row][:, 0]).max() <= 1e-9 * np.abs(gg[:nf][row][:, 1]).max()


def test_gamma_face_edge_corner_3d():
    # the evaluation point is moved off the wall by h^2 taken as a LENGTH (fluid_equations.hpp:187):
    # a small tank keeps that a small fraction of h
    case = cases.dam_break_3d(5, H=0.05, tank=(2.6, 2.4, 2.4), wall_ratio=0.93, containment_margin=0.5)
    s = oracle_lib.OracleSolver(3)
    oracle_lib.load_case(s, case)
    s.initialize()
    gamma = s.download("gamma")[case.n_fluid:]
    rx = case.r[case.n_fluid:]
    ext = np.asarray(case.meta["tank"])
    radius = 2.0 * case.h
    on = (np.abs(rx) < 1e-12) | (np.abs(rx - ext) < 1e-12)
    nwall = on.sum(axis=1)
    dist_other = np.where(on, np.inf, np.minimum(rx, ext - rx)).min(axis=1)  # distance to the nearest OTHER wall plane
    face = (nwall == 1) & (dist_other > radius)
    edge = (nwall == 2) & (dist_other > radius)
    corner = nwall == 3
    assert face.sum() > 20 and edge.sum() > 10 and corner.sum() == 8
    assert np.abs(gamma[face] - 0.5).max() <= 0.03
    assert np.abs(gamma[edge] - 0.25).max() <= 0.03
    assert np.abs(gamma[corner] - 0.125).max() <= 0.03


def still_tank_2d(n_col, H=0.6):
    """A closed 2-D tank filled from wall to wall up to H with the hydrostatic (Tait) density."""
    dr = H / n_col
    PW, PH = dr * (n_col + 1), 2.0 * H
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    verts, faces = cases.tessellate_2d(np.array([[0.0, PH], [PW, PH], [PW, 0.0], [0.0, 0.0]]), np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.uint64), dr)
    cverts = np.array([[0.0, 0.0], [PW, 0.0], [PW, PH], [0.0, PH]])
    cfaces = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.uint64)
    ii, jj = np.meshgrid(np.arange(n_col), np.arange(n_col), indexing="ij")
    rf = dr * np.stack([ii.ravel() + 1.0, jj.ravel() + 1.0], axis=1)
    r = np.concatenate([rf, verts])
    nf, nx = len(rf), len(verts)
    m, rho = np.full(nf + nx, rho0 * dr * dr), np.full(nf + nx, rho0)
    p = rho0 * g * (dr * (n_col + 0.5) - rf[:, 1])  # free surface half a spacing above the top row
    rho[:nf] = rho0 * (1.0 + 7.0 * p / (rho0 * cs0**2)) ** (1.0 / 7.0)
    return cases.Case(2, nf, nx, r, m, rho, verts, faces, cverts, cfaces, g, 1e-3, cs0, rho0, 7.0, 2 * dr, dr, H, {"name": "still_tank_2d", "tank": (PW, PH)})


def test_hydrostatic_column_is_in_equilibrium():
    """A tank filled from wall to wall with the hydrostatic density profile: in the interior
    the pair sums of the pressure term balance gravity to 1 % (the dam-break column itself is
    NOT in equilibrium - its free side accelerates at t = 0)."""
    case = still_tank_2d(30)
    s = oracle_lib.OracleSolver(2)
    oracle_lib.load_case(s, case)
    s.initialize()
    s.rhs_only()
    nf = case.n_fluid
    rf = case.r[:nf]
    radius = 2.0 * case.h
    inner = (rf[:, 0] > radius) & (rf[:, 0] < case.meta["tank"][0] - radius) & (rf[:, 1] > radius) & (rf[:, 1] < case.H - 1.5 * radius)
    assert inner.sum() > 300
    dv = s.download("dv_dt")[:nf][inner]
    assert np.abs(dv).max() <= 0.02 * case.g
    # without the pressure term the same particles would fall freely
    assert np.abs(s.download("gamma")[:nf][inner] - 1.0).max() == 0.0
    drho = s.download("drho_dt")[:nf][inner]
    assert np.abs(drho).max() * 1e-4 <= 1e-4 * case.rho0  # over a time step (1e-4 s) the density moves by < 1e-4 rho0


@pytest.mark.parametrize("dim", [2, 3])
def test_symmetric_block_coloured_pair_loops_equal_the_gather_form(dim):
    """The CPU baseline of bench.py runs the pair sums as the reference does - once per unordered
    pair, both particles updated, race-free by a partition into blocks (oracle_sim.h: for_each_pair;
    particle_mesh.hpp:165-241) - while the parity oracle keeps the gather form. Same sums in another
    order: right-hand sides, shifting sums and a few whole steps must agree to rounding."""
    import oracle_lib
    from titsolver_b200 import cases

    case = cases.dam_break_2d(24) if dim == 2 else cases.dam_break_3d(8, wall_ratio=0.93, jitter=0.1)
    rng = np.random.default_rng(11)
    v = np.zeros((case.n, dim))
    v[: case.n_fluid] = 0.3 * rng.standard_normal((case.n_fluid, dim))
    out = []
    for sym in (False, True):
        c = oracle_lib.OracleSolver(dim, 4, 0, 3)
        oracle_lib.load_case(c, case)
        c.upload("v", v)
        c.set_symmetric(sym)
        c.initialize()
        c.rhs_only()
        res = {"rhs_" + f: c.download(f) for f in ("drho_dt", "dv_dt")}
        c.step(3)
        res.update({f: c.download(f) for f in ("r", "v", "rho", "N", "L", "grad_v", "grad_rho", "dr", "phi")})
        out.append(res)
    nf = case.n_fluid
    for f in out[0]:
        a, b = out[0][f][:nf], out[1][f][:nf]
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(a).max(), 1e-300), f
