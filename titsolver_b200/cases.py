"""Synthetic dam-break case setup (host side, numpy; setup only — not the hot path).

Mirrors the reference driver /root/reference/source/titwcsph/wcsph.cpp:37-142
(constants, tank surface + 2-D tessellation, lattice fill, hydrostatic density)
and /root/reference/source/tit/geom/tessellation.hpp:30-63 (2-D tessellate), :74-163 (3-D).
The 3-D case has no reference counterpart: it uses the
same constants with a structured wall mesh of vertex spacing `wall_dr`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

KERNELS = {
    "CubicSpline": 0,
    "QuarticSpline": 1,
    "QuinticSpline": 2,
    "QuarticWendland": 3,
    "SixthOrderWendland": 4,
    "EighthOrderWendland": 5,
}
EOS = {"tait": 0, "linear_tait": 1}
INTEGRATORS = {"symplectic_euler": 0, "velocity_verlet": 1, "ssprk2": 2, "ssprk3": 3}


@dataclass
class Case:
    dim: int
    n_fluid: int
    n_fixed: int
    r: np.ndarray  # (n, dim) fluid first, then fixed (vertex order)
    m: np.ndarray
    rho: np.ndarray
    verts: np.ndarray  # boundary-integral surface (normals point into the fluid)
    faces: np.ndarray  # (nf, dim) uint64 vertex indices
    cverts: np.ndarray  # containment surface (winding number +1 inside)
    cfaces: np.ndarray
    g: float
    mu: float
    cs0: float
    rho0: float
    xi: float
    h: float
    dr: float
    H: float
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.n_fluid + self.n_fixed


def tessellate_2d(verts: np.ndarray, faces: np.ndarray, d_max: float):
    """geom/tessellation.hpp:30-63: split each segment into ceil(len/d_max) parts."""
    out_v = [tuple(v) for v in verts]
    out_f = []
    for a_i, b_i in faces:
        a = np.asarray(verts[a_i], dtype=np.float64)
        b = np.asarray(verts[b_i], dtype=np.float64)
        ba = b - a
        d = math.sqrt(ba[1] * ba[1] + ba[0] * ba[0])  # norm(cross(ba))
        n = max(1, int(math.ceil(d / d_max)))
        prev = int(a_i)
        for i in range(1, n):
            t = float(i) / float(n)
            out_v.append(tuple(a + t * ba))
            vi = len(out_v) - 1
            out_f.append((prev, vi))
            prev = vi
        out_f.append((prev, int(b_i)))
    return np.asarray(out_v, dtype=np.float64), np.asarray(out_f, dtype=np.uint64)


def tessellate_3d(verts: np.ndarray, faces: np.ndarray, d_max: float):
    """geom/tessellation.hpp:74-163: red refinement of a triangle surface until no
    edge is longer than `d_max`. Per sweep every too-long edge gets ONE midpoint
    (vertices are appended in the order the edges are met: faces in order, edges
    ab, bc, ca), then every triangle is cut along its midpoints. With the
    triangle rotated so that its first split edge is (p, q), the children are
        3 midpoints   (p, m_pq, m_sp) (m_pq, q, m_qs) (m_sp, m_qs, s) (m_pq, m_qs, m_sp)
        m_pq + m_qs   (p, m_pq, s) (m_pq, m_qs, s) (m_pq, q, m_qs)
        m_pq          (p, m_pq, s) (m_pq, q, s)
    which reproduces the vertex and face numbering of the reference's tests
    (geom/tessellation.test.cpp:64-215)."""
    V = [tuple(float(x) for x in v) for v in verts]
    F = [tuple(int(i) for i in f) for f in faces]
    d2_max = float(d_max) * float(d_max)
    while True:
        mid = {}

        def visit(i, j):
            key = (i, j) if i < j else (j, i)
            if key in mid:
                return
            a, b = V[i], V[j]
            if (b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2 + (b[2] - a[2]) ** 2 <= d2_max:
                return
            V.append(((a[0] + b[0]) / 2, (a[1] + b[1]) / 2, (a[2] + b[2]) / 2))
            mid[key] = len(V) - 1

        for a, b, c in F:
            visit(a, b); visit(b, c); visit(c, a)
        if not mid:
            break
        out = []
        for tri in F:
            m = [mid.get((min(tri[k], tri[(k + 1) % 3]), max(tri[k], tri[(k + 1) % 3]))) for k in range(3)]
            n_split = sum(x is not None for x in m)
            if n_split == 0:
                out.append(tri)
                continue
            # rotation: the first split edge in cyclic order whose predecessor is not split (any for 3)
            r = 0 if n_split == 3 else next(k for k in range(3) if m[k] is not None and m[(k + 2) % 3] is None)
            p, q, s = tri[r], tri[(r + 1) % 3], tri[(r + 2) % 3]
            m_pq, m_qs, m_sp = m[r], m[(r + 1) % 3], m[(r + 2) % 3]
            if n_split == 3:
                out += [(p, m_pq, m_sp), (m_pq, q, m_qs), (m_sp, m_qs, s), (m_pq, m_qs, m_sp)]
            elif n_split == 2:
                out += [(p, m_pq, s), (m_pq, m_qs, s), (m_pq, q, m_qs)]
            else:
                out += [(p, m_pq, s), (m_pq, q, s)]
        F = out
    return np.asarray(V, dtype=np.float64), np.asarray(F, dtype=np.uint64)


def dam_break_2d(n_col: int = 80, H: float = 0.6) -> Case:
    """wcsph.cpp:37-142 with `dr = H / n_col` (the reference uses n_col = 80)."""
    L = 2 * H
    PW, PH = 5.366 * H, 4.0 * H
    dr = H / float(n_col)
    WM, WN = int(round(L / dr)), int(round(H / dr))
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    h0 = 2.0 * dr
    m0 = rho0 * dr**2
    mu = 0.001
    # Clockwise: Segment::normal() = normalize((ba.y, -ba.x)) points inwards.
    dv = np.array([[0.0, PH], [PW, PH], [PW, 0.0], [0.0, 0.0]])
    df = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.uint64)
    verts, faces = tessellate_2d(dv, df, dr)
    cverts = np.array([[0.0, 0.0], [PW, 0.0], [PW, PH], [0.0, PH]])
    cfaces = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.uint64)
    ii, jj = np.meshgrid(np.arange(WM), np.arange(WN), indexing="ij")
    rf = dr * np.stack([ii.ravel() + 1.0, jj.ravel() + 1.0], axis=1)
    r = np.concatenate([rf, verts], axis=0)
    nf, nx = rf.shape[0], verts.shape[0]
    m = np.full(nf + nx, m0)
    rho = np.full(nf + nx, rho0)
    # Hydrostatic density from the Poisson series (wcsph.cpp:119-142).
    x, y = rf[:, 0], rf[:, 1]
    p = rho0 * g * (H - y)
    for k in range(1, 100, 2):
        kpi = float(k) * math.pi
        p = p - 8 * rho0 * g * H / kpi**2 * (np.exp(kpi * (x - L) / (2 * H)) * np.cos(kpi * y / (2 * H)))
    rho[:nf] = rho0 + p / cs0**2
    return Case(2, nf, nx, r, m, rho, verts, faces, cverts, cfaces, g, mu, cs0, rho0, 7.0, h0, dr, H,
                {"name": f"dam_break_2d_{WM}x{WN}", "tank": (PW, PH)})


def _box_wall_mesh(ext, n_cells, inward=True, cell_range=None):
    """Structured triangulation of the 6 walls of [0,ext]; two right triangles per
    quad; vertices shared along edges. Triangle normals (cross(ba, ca)) point
    into the box when `inward`. Vectorised: the 10 M-particle case has 1.8 M
    wall vertices. `cell_range = (axis, c0, c1)` keeps only the quads whose cell
    index along `axis` lies in [c0, c1) (and the end walls normal to `axis` only
    if the range reaches them): the part of the mesh one slab of a domain
    decomposition needs, with coordinates identical to the full mesh."""
    n = np.asarray(n_cells, dtype=np.int64)
    tris = []  # (nt, 3, 3) integer lattice coordinates of the triangle corners
    r_axis, c0, c1 = cell_range if cell_range is not None else (-1, 0, 0)

    def rng(a):
        return np.arange(max(c0, 0), min(c1, n[a])) if a == r_axis else np.arange(n[a])

    def wall(axis, level, u_axis, v_axis, flip):
        if axis == r_axis and not ((level == 0 and c0 <= 0) or (level == n[axis] and c1 >= n[axis])):
            return
        # Quad corners p00 -> p10 (+u) -> p11 -> p01 (+v); normal = e_u x e_v.
        iu, iv = np.meshgrid(rng(u_axis), rng(v_axis), indexing="ij")
        iu, iv = iu.ravel(), iv.ravel()

        def P(du, dv):
            c = np.zeros((iu.size, 3), dtype=np.int64)
            c[:, axis] = level
            c[:, u_axis] = iu + du
            c[:, v_axis] = iv + dv
            return c

        p00, p10, p11, p01 = P(0, 0), P(1, 0), P(1, 1), P(0, 1)
        if flip:
            t1, t2 = (p00, p11, p10), (p00, p01, p11)
        else:
            t1, t2 = (p00, p10, p11), (p00, p11, p01)
        # interleave the two triangles of each quad
        q = np.stack([np.stack(t1, axis=1), np.stack(t2, axis=1)], axis=1)  # (nq, 2, 3, 3)
        tris.append(q.reshape(-1, 3, 3))

    wall(2, 0, 0, 1, not inward)      # z = 0, normal +z
    wall(2, n[2], 0, 1, inward)       # z = ez
    wall(1, 0, 2, 0, not inward)      # y = 0, normal +y (e_z x e_x)
    wall(1, n[1], 2, 0, inward)
    wall(0, 0, 1, 2, not inward)      # x = 0, normal +x (e_y x e_z)
    wall(0, n[0], 1, 2, inward)
    if not tris:
        return np.zeros((0, 3)), np.zeros((0, 3), np.uint64)
    t = np.concatenate(tris, axis=0)
    key = (t[..., 0] * (n[1] + 1) + t[..., 1]) * (n[2] + 1) + t[..., 2]
    ukey, inv = np.unique(key.ravel(), return_inverse=True)
    faces = inv.reshape(-1, 3).astype(np.uint64)
    k = ukey % (n[2] + 1)
    j = (ukey // (n[2] + 1)) % (n[1] + 1)
    i = ukey // ((n[2] + 1) * (n[1] + 1))
    verts = np.stack([ext[0] * i / n[0], ext[1] * j / n[1], ext[2] * k / n[2]], axis=1).astype(np.float64)
    return verts, faces


def dam_break_3d(n_col: int = 16, H: float = 0.6, wall_ratio: float = 1.0, tank=(5.366, 4.0, 1.0),
                 jitter: float = 0.0, seed: int = 123, containment_margin: float = 0.5) -> Case:
    """3-D dam break: column 2H x H x H (x, y up, z) of n_col particles per H in a
    closed box `tank`*H. Fixed particles sit on the wall-mesh vertices (as in the
    2-D reference case); the wall vertex spacing is about `wall_ratio * dr`.

    There is no 3-D case in the reference, so two set-up choices are ours:
      * the containment box is grown by `containment_margin * dr` so that the
        wall particles are strictly inside it. On the surface itself the
        reference's test `winding > 0.5` is decided by rounding noise
        (geom/winding/exact_winding.hpp:41-43), and near tank edges the two
        outcomes give different gamma (fluid_equations.hpp:183-191);
      * `jitter` (in units of dr, seeded) moves the fluid particles off the
        lattice and `wall_ratio != 1` makes the wall mesh incommensurate with
        the support radius 4 dr. The reference's 3-D triangle integral skips
        boundary pieces shorter than `tiny` (sph/kernel.hpp:371-393), so faces
        that touch the support sphere exactly are evaluated at rounding-noise
        level either way; parity tests use generic positions.
    """
    dr = H / float(n_col)
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    h0 = 2.0 * dr
    m0 = rho0 * dr**3
    mu = 0.001
    ext = (tank[0] * H, tank[1] * H, tank[2] * H)
    wd = wall_ratio * dr
    ncell = tuple(max(1, int(math.ceil(e / wd - 1e-9))) for e in ext)
    verts, faces = _box_wall_mesh(ext, ncell, inward=True)
    # Containment: a 12-triangle box, outward normals => winding +1 inside.
    cverts, cfaces = _box_wall_mesh(ext, (1, 1, 1), inward=False)
    if containment_margin:
        mg = containment_margin * dr
        cverts = np.where(cverts > 0.0, cverts + mg, cverts - mg)
    WM, WN, WK = 2 * n_col, n_col, int(round(tank[2] * n_col)) - 1
    ii, jj, kk = np.meshgrid(np.arange(WM), np.arange(WN), np.arange(WK), indexing="ij")
    rf = dr * np.stack([ii.ravel() + 1.0, jj.ravel() + 1.0, kk.ravel() + 1.0], axis=1)
    if jitter:
        rng = np.random.default_rng(seed)
        rf = rf + rng.uniform(-jitter, jitter, size=rf.shape) * dr
    r = np.concatenate([rf, verts], axis=0)
    nf, nx = rf.shape[0], verts.shape[0]
    m = np.full(nf + nx, m0)
    rho = np.full(nf + nx, rho0)
    rho[:nf] = rho0 + rho0 * g * (H - rf[:, 1]) / cs0**2
    return Case(3, nf, nx, r, m, rho, verts, faces, cverts, cfaces, g, mu, cs0, rho0, 7.0, h0, dr, H,
                {"name": f"dam_break_3d_{WM}x{WN}x{WK}", "tank": ext, "wall_ratio": wall_ratio, "jitter": jitter})


def dam_break_3d_counts(n_col: int, tank=(5.366, 4.0, 1.0)):
    """(n_fluid, n_fixed) of `dam_break_3d(n_col, tank=tank)` without building it."""
    ncell = tuple(max(1, int(math.ceil(t * n_col - 1e-9))) for t in tank)
    n_fixed = int(np.prod([c + 1 for c in ncell]) - np.prod([c - 1 for c in ncell]))
    return 2 * n_col * n_col * (int(round(tank[2] * n_col)) - 1), n_fixed


def fluid_block_3d(nx: int, ny: int, nz: int, H: float = 0.6, n_col: int = 171) -> Case:
    """A block of nx x ny x nz fluid particles at the spacing of `dam_break_3d(n_col)` with no
    walls at all (gamma = 1 everywhere): the bulk part of the dam break, for timing the
    per-fluid-particle cost of a CPU implementation apart from its wall integrals."""
    dr = H / float(n_col)
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    rf = dr * np.stack([ii.ravel() + 1.0, jj.ravel() + 1.0, kk.ravel() + 1.0], axis=1)
    nf = rf.shape[0]
    big = 10.0 * dr * max(nx, ny, nz) + 1.0
    cverts, cfaces = _box_wall_mesh((2 * big,) * 3, (1, 1, 1), inward=False)
    cverts = cverts - big
    rho = rho0 + rho0 * g * (dr * ny - rf[:, 1]) / cs0**2
    e = np.zeros((0, 3))
    return Case(3, nf, 0, rf, np.full(nf, rho0 * dr**3), rho, e, e.astype(np.uint64), cverts, cfaces, g, 0.001, cs0, rho0, 7.0, 2.0 * dr, dr, H,
                {"name": f"fluid_block_3d_{nx}x{ny}x{nz}", "tank": (2 * big,) * 3})


def tetra_tank_3d(n_side: int = 10, L: float = 1.0, wall_ratio: float = 0.93, jitter: float = 0.1, seed: int = 7, fill: float = 0.55) -> Case:
    """Fluid in a tetrahedral tank (0,0,0), (L,0,0), (0,L,0), (0,0,L): a wall surface
    with slanted, red-refined triangles (`tessellate_3d`, the reference's 3-D set-up
    path, geom/tessellation.hpp:74-163) instead of the axis-aligned structured walls
    of `dam_break_3d`. The fluid fills the part of the tank below the plane
    x + y + z = fill * L on a jittered lattice of spacing dr = L / n_side, at least
    half a spacing away from the walls. Parity cases only (generic positions)."""
    dr = L / float(n_side)
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * L)
    h0 = 2.0 * dr
    tet = np.array([[0.0, 0.0, 0.0], [L, 0.0, 0.0], [0.0, L, 0.0], [0.0, 0.0, L]])
    out_faces = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.uint64)  # outward normals
    verts, faces = tessellate_3d(tet, out_faces[:, [0, 2, 1]], wall_ratio * dr)    # normals into the fluid
    # Containment: the tetrahedron blown up about its centroid so that the wall particles are inside.
    ctr = tet.mean(axis=0)
    cverts, cfaces = ctr + (tet - ctr) * (1.0 + 3.0 * dr / L), out_faces
    k = np.arange(n_side)
    ii, jj, kk = np.meshgrid(k, k, k, indexing="ij")
    rf = dr * (np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1) + 0.75)
    rng = np.random.default_rng(seed)
    rf = rf + rng.uniform(-jitter, jitter, size=rf.shape) * dr
    s = rf.sum(axis=1)
    rf = rf[(rf.min(axis=1) >= 0.5 * dr) & (s <= fill * L)]
    nf, nx = rf.shape[0], verts.shape[0]
    r = np.concatenate([rf, verts], axis=0)
    return Case(dim=3, n_fluid=nf, n_fixed=nx, r=r, m=np.full(nf + nx, rho0 * dr**3), rho=np.full(nf + nx, rho0), verts=verts, faces=faces,
                cverts=cverts, cfaces=cfaces, g=g, mu=0.001, cs0=cs0, rho0=rho0, xi=7.0, h=h0, dr=dr, H=L,
                meta={"name": f"tetra_tank_3d_{n_side}", "wall_ratio": wall_ratio, "jitter": jitter})


def dam_break_3d_slab(n_col: int, world: int, rank: int, H: float = 0.6, tank=(5.366, 4.0, 1.0), halo_cells: int = 18, mode: str = "weak_z"):
    """Rank-local part of a decomposed 3-D dam break.

    mode "weak_z"   (weak scaling) the tank of `dam_break_3d` made `world` times deeper
                    along z (the flow is z-invariant, so the slabs stay balanced), cut into
                    `world` slabs along z;
    mode "strong_x" (strong scaling) the tank of `dam_break_3d` itself, the fluid column cut
                    into `world` slabs of equally many lattice planes along x; the last
                    rank also holds the dry part of the tank.

    Returns (case, edges, axis): `case` holds the fluid particles this rank owns
    (`case.meta["gid"]` = their global lattice indices), the wall vertices /
    faces within `halo_cells` wall cells of the slab (fixed particles = those
    vertices) and the global containment box. Coordinates are bit-identical to
    those of the global case (`dam_break_3d(n_col, tank=...)`).
    """
    dr = H / float(n_col)
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    h0 = 2.0 * dr
    m0 = rho0 * dr**3
    mu = 0.001
    if mode == "weak_z":
        axis, tz = 2, tank[2] * world
    elif mode == "strong_x":
        axis, tz = 0, tank[2]
    else:
        raise ValueError(mode)
    ext = (tank[0] * H, tank[1] * H, tz * H)
    ncell = tuple(max(1, int(math.ceil(e / dr - 1e-9))) for e in ext)
    WM, WN, WK = 2 * n_col, n_col, int(round(tz * n_col)) - 1
    dims = (WM, WN, WK)
    if mode == "weak_z":
        per = int(round(tank[2] * n_col))  # lattice planes per slab: plane k (z = dr (k + 1)) belongs to rank (k + 1) // per
        cuts = [0] + [per * j - 1 for j in range(1, world)] + [WK]
    else:
        cuts = [(WM * j) // world for j in range(world + 1)]
    # slab edges half-way between lattice planes (plane i sits at dr (i + 1))
    edges = [-math.inf] + [dr * (cuts[j] + 0.5) for j in range(1, world)] + [math.inf]
    k0, k1 = cuts[rank], cuts[rank + 1]
    rng = [np.arange(d) for d in dims]
    rng[axis] = np.arange(k0, k1)
    ii, jj, kk = np.meshgrid(*rng, indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    rf = dr * np.stack([ii + 1.0, jj + 1.0, kk + 1.0], axis=1)
    gid = (ii.astype(np.int64) * WN + jj) * WK + kk
    c0 = (k0 + 1) - halo_cells if rank > 0 else 0
    c1 = (k1 + 1) + halo_cells if rank < world - 1 else ncell[axis]
    verts, faces = _box_wall_mesh(ext, ncell, inward=True, cell_range=(axis, c0, c1))
    cverts, cfaces = _box_wall_mesh(ext, (1, 1, 1), inward=False)
    mg = 0.5 * dr
    cverts = np.where(cverts > 0.0, cverts + mg, cverts - mg)
    r = np.concatenate([rf, verts], axis=0)
    nf, nx = rf.shape[0], verts.shape[0]
    m = np.full(nf + nx, m0)
    rho = np.full(nf + nx, rho0)
    rho[:nf] = rho0 + rho0 * g * (H - rf[:, 1]) / cs0**2
    n_fixed_global = int(np.prod([c + 1 for c in ncell]) - np.prod([c - 1 for c in ncell]))
    case = Case(3, nf, nx, r, m, rho, verts, faces, cverts, cfaces, g, mu, cs0, rho0, 7.0, h0, dr, H,
                {"name": f"dam_break_3d_{WM}x{WN}x{WK}_{mode}_slab{rank}of{world}", "tank": ext, "gid": gid,
                 "n_fluid_global": WM * WN * WK, "n_fixed_global": n_fixed_global})
    return case, edges, axis
