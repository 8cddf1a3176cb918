"""Synthetic dam-break case setup (host side, numpy; setup only — not the hot path).

Mirrors the reference driver /root/reference/source/titwcsph/wcsph.cpp:37-142
(constants, tank surface + 2-D tessellation, lattice fill, hydrostatic density)
and /root/reference/source/tit/geom/tessellation.hpp:30-63 (2-D tessellate).
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


def _box_wall_mesh(ext, n_cells, inward=True):
    """Structured triangulation of the 6 walls of [0,ext]; two right triangles per
    quad; vertices shared along edges. Triangle normals (cross(ba, ca)) point
    into the box when `inward`."""
    nx, ny, nz = n_cells
    ex, ey, ez = ext
    vid = {}
    verts = []

    def V(i, j, k):
        key = (i, j, k)
        if key not in vid:
            vid[key] = len(verts)
            verts.append((ex * i / nx, ey * j / ny, ez * k / nz))
        return vid[key]

    faces = []

    def quad(p00, p10, p11, p01, flip):
        # p00->p10->p11->p01 counter-clockwise seen from the side the normal points to.
        if flip:
            faces.append((p00, p11, p10))
            faces.append((p00, p01, p11))
        else:
            faces.append((p00, p10, p11))
            faces.append((p00, p11, p01))

    # z = 0 (normal +z) and z = ez (normal -z)
    for i in range(nx):
        for j in range(ny):
            quad(V(i, j, 0), V(i + 1, j, 0), V(i + 1, j + 1, 0), V(i, j + 1, 0), not inward)
            quad(V(i, j, nz), V(i + 1, j, nz), V(i + 1, j + 1, nz), V(i, j + 1, nz), inward)
    # y = 0 (normal +y): CCW seen from +y is (x, z) order z then x
    for i in range(nx):
        for k in range(nz):
            quad(V(i, 0, k), V(i, 0, k + 1), V(i + 1, 0, k + 1), V(i + 1, 0, k), not inward)
            quad(V(i, ny, k), V(i, ny, k + 1), V(i + 1, ny, k + 1), V(i + 1, ny, k), inward)
    # x = 0 (normal +x): CCW seen from +x is (y, z)
    for j in range(ny):
        for k in range(nz):
            quad(V(0, j, k), V(0, j + 1, k), V(0, j + 1, k + 1), V(0, j, k + 1), not inward)
            quad(V(nx, j, k), V(nx, j + 1, k), V(nx, j + 1, k + 1), V(nx, j, k + 1), inward)
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.uint64)


def dam_break_3d(n_col: int = 16, H: float = 0.6, wall_stride: int = 1, tank=(5.366, 4.0, 1.0)) -> Case:
    """3-D dam break: column 2H x H x H (x, y up, z) of n_col particles per H in a
    closed box `tank`*H. Wall vertex spacing = wall_stride * dr (fixed particles
    sit on the wall vertices, as in the 2-D reference case)."""
    dr = H / float(n_col)
    g, rho0 = 9.81, 1000.0
    cs0 = 20 * math.sqrt(g * H)
    h0 = 2.0 * dr
    m0 = rho0 * dr**3
    mu = 0.001
    ext = (tank[0] * H, tank[1] * H, tank[2] * H)
    wd = wall_stride * dr
    ncell = tuple(max(1, int(math.ceil(e / wd - 1e-9))) for e in ext)
    verts, faces = _box_wall_mesh(ext, ncell, inward=True)
    # Containment: the 12-triangle box, outward normals => winding +1 inside.
    cverts, cfaces = _box_wall_mesh(ext, (1, 1, 1), inward=False)
    WM, WN, WK = 2 * n_col, n_col, int(round(tank[2] * n_col)) - 1
    ii, jj, kk = np.meshgrid(np.arange(WM), np.arange(WN), np.arange(WK), indexing="ij")
    rf = dr * np.stack([ii.ravel() + 1.0, jj.ravel() + 1.0, kk.ravel() + 1.0], axis=1)
    r = np.concatenate([rf, verts], axis=0)
    nf, nx = rf.shape[0], verts.shape[0]
    m = np.full(nf + nx, m0)
    rho = np.full(nf + nx, rho0)
    rho[:nf] = rho0 + rho0 * g * (H - rf[:, 1]) / cs0**2
    return Case(3, nf, nx, r, m, rho, verts, faces, cverts, cfaces, g, mu, cs0, rho0, 7.0, h0, dr, H,
                {"name": f"dam_break_3d_{WM}x{WN}x{WK}", "tank": ext, "wall_stride": wall_stride})
