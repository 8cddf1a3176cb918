"""Pins the oracle's geometry layer against the reference's own unit tests.

Restates (all paths under /root/reference/source/tit/geom/):
  search.test.cpp:130-163       grid search == brute force as sets; std::mt19937{123}
                                clouds (reproduced bit for bit), radii {.01,.1,.5,1},
                                cell hints {0.5, 5} x radius, 2D / 3D / empty
  grid.test.cpp:62-160          cell extents / counts, cells_intersecting
  segment.test.cpp:70-158       winding number, clamp (incl. degenerate), intersects
  triangle.test.cpp:95-258      winding number, clamp (incl. degenerate), intersects
  face_search.test.cpp:108-160  face search == brute force on tessellated surfaces
  winding.test.cpp:62-100       exact winding of a square / tetrahedron
"""
import numpy as np
import pytest

import oracle_lib as ol
from titsolver_b200 import cases

TINY = float(np.cbrt(np.finfo(float).eps))


def approx(a, b, eps=TINY):
    d = np.atleast_1d(np.asarray(a, float) - np.asarray(b, float))
    return float(np.dot(d, d)) <= eps * eps


class StdMt19937Uniform:
    """std::mt19937{seed} + std::uniform_real_distribution<double>{0, 1} as libstdc++
    evaluates it: generate_canonical<double, 53> draws two 32-bit words,
    (x0 + x1 * 2^32) / 2^64."""

    def __init__(self, seed):
        rs = np.random.RandomState(seed)  # init_genrand(seed), same as std::mt19937{seed}
        self.bg = np.random.MT19937()
        st = rs.get_state()
        self.bg.state = {"bit_generator": "MT19937", "state": {"key": st[1], "pos": st[2]}}

    def __call__(self):
        x0, x1 = (int(v) for v in self.bg.random_raw(2))
        r = (float(x0) + float(x1) * 4294967296.0) / 18446744073709551616.0
        return r if r < 1.0 else np.nextafter(1.0, 0.0)


def test_std_mt19937_reproduction():
    # 10000th output of std::mt19937{} (seed 5489) is 4123659995 (C++ standard, [rand.predef]).
    g = StdMt19937Uniform(5489)
    assert int(g.bg.random_raw(10000)[-1]) == 4123659995


def reference_cloud(dim, n=200):
    gen = StdMt19937Uniform(123)
    return np.array([[gen() for _ in range(dim)] for _ in range(n)])


def brute_neighbors(pts, radius):
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    return [set(np.nonzero(row < radius * radius)[0].tolist()) | {i} for i, row in enumerate(d2)]


def csr_rows(off, cols):
    return [set(cols[int(off[i]):int(off[i + 1])].tolist()) for i in range(len(off) - 1)]


def oracle_search(pts, radius, hint, dim, verts=None, faces=None, want_faces=False):
    s = ol.OracleSolver(dim, 0)  # cubic spline: support radius = 2 h
    s.set_params(9.81, 1e-3, 10.0, 1000.0, 7.0, radius / 2.0, hint, hint)
    ev, ef = np.zeros((0, dim)), np.zeros((0, dim), np.uint64)
    s.set_surface(ev if verts is None else verts, ef if faces is None else faces, ev, ef)
    s.set_particles(len(pts), 0)
    if len(pts):
        s.upload("r", pts)
    return s.face_neighbors() if want_faces else s.neighbors()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("radius", [0.01, 0.1, 0.5, 1.0])
@pytest.mark.parametrize("scale", [0.5, 5.0])
def test_grid_search_matches_brute_force(dim, radius, scale):
    pts = reference_cloud(dim)
    assert 0.0 <= pts.min() and pts.max() < 1.0
    off, cols = oracle_search(pts, radius, scale * radius, dim)
    assert csr_rows(off, cols) == brute_neighbors(pts, radius)
    # rows come out sorted (sph/particle_mesh.hpp:146)
    for i in range(len(pts)):
        row = cols[int(off[i]):int(off[i + 1])]
        assert np.all(np.diff(row.astype(np.int64)) > 0)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("radius", [0.01, 0.1, 0.5, 1.0])
def test_kd_tree_search_matches_brute_force(dim, radius):
    """geom/search.test.cpp:122-125: the K-d tree index on the same clouds."""
    pts = reference_cloud(dim)
    off, cols = ol.kdtree_neighbors(pts, radius)
    assert csr_rows(off, cols) == brute_neighbors(pts, radius)
    # ... and identical, row by row, to the grid index (both feed ParticleMesh::search_)
    goff, gcols = oracle_search(pts, radius, radius, dim)
    assert np.array_equal(off, goff) and np.array_equal(cols, gcols)


def test_kd_tree_search_empty_single_and_duplicates():
    off, cols = ol.kdtree_neighbors(np.zeros((0, 2)), 0.1)
    assert len(cols) == 0 and list(off) == [0]
    off, cols = ol.kdtree_neighbors(np.array([[0.3, 0.4, 0.5]]), 0.1)
    assert list(off) == [0, 1] and list(cols) == [0]
    # Coincident points, and points exactly one radius apart. The point test is inclusive
    # (bsphere.hpp:52-53) but the descent to the far side of a splitting plane is strict
    # (kd_tree_search.hpp:116-120: pow2(delta) < pow2(radius)), so a point exactly one radius
    # away along the split axis may or may not be reported, depending on where the tree was
    # split: every row lies between the strict and the inclusive brute-force answer. (The
    # grid index, and the GPU hash behind either option, always give the inclusive answer.)
    pts = np.array([[0.0, 0.0], [0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [0.0, 0.5], [0.0, 0.0]])
    off, cols = ol.kdtree_neighbors(pts, 0.5)
    rows = csr_rows(off, cols)
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    for i, row in enumerate(rows):
        strict = set(np.nonzero(d2[i] < 0.25)[0].tolist())
        inclusive = set(np.nonzero(d2[i] <= 0.25)[0].tolist())
        assert strict <= row <= inclusive, (i, row)
    assert rows[0] == {0, 1, 2, 4, 5} and rows[3] == {2, 3}
    goff, gcols = oracle_search(pts, 0.5, 0.5, 2)
    assert csr_rows(goff, gcols) == [set(np.nonzero(d2[i] <= 0.25)[0].tolist()) for i in range(len(pts))]


def test_grid_search_empty():
    off, cols = oracle_search(np.zeros((0, 2)), 0.1, 0.05, 2)
    assert len(cols) == 0 and list(off) == [0]


# ---- grid.test.cpp
def grid(lo, hi, hint):
    dim = len(lo)
    num, ext = np.zeros(dim, np.uint64), np.zeros(dim)
    assert ol.load().orc_grid(dim, ol._dp(ol._arr(lo)), ol._dp(ol._arr(hi)), hint, ol._u64p(num), ol._dp(ext)) == 0
    return num.tolist(), ext.tolist()


def test_grid_cell_extents():
    """grid.test.cpp:62-81 — uniform hints (GridIndex only ever passes one, grid_search.hpp:55-57)."""
    assert grid([0, 0], [8, 8], 4.0) == ([2, 2], [4.0, 4.0])          # divides
    assert grid([0, 0], [8, 8], 2.0) == ([4, 4], [2.0, 2.0])
    assert grid([0, 0], [8, 8], 5.0) == ([2, 2], [4.0, 4.0])          # does not divide: ceil, then stretch
    assert grid([0, 0], [8, 8], 2.5) == ([4, 4], [2.0, 2.0])
    assert grid([0, 0], [4, 4], 7.0) == ([1, 1], [4.0, 4.0])          # too large
    assert grid([0, 0, 0], [8, 4, 2], 2.0) == ([4, 2, 1], [2.0, 2.0, 2.0])


def cells_intersecting(lo, hi, hint, qlo, qhi):
    dim = len(lo)
    a, b = np.zeros(dim, np.uint64), np.zeros(dim, np.uint64)
    rc = ol.load().orc_grid_cells_intersecting(dim, ol._dp(ol._arr(lo)), ol._dp(ol._arr(hi)), hint, ol._dp(ol._arr(qlo)), ol._dp(ol._arr(qhi)), ol._u64p(a), ol._u64p(b))
    return None if rc else (a.tolist(), b.tolist())


def test_grid_cells_intersecting():
    """grid.test.cpp:139-160: 8x8 grid over [0,8]^2, inclusive cell ranges."""
    assert cells_intersecting([0, 0], [8, 8], 1.0, [3, 3], [5, 5]) == ([3, 3], [5, 5])   # full
    assert cells_intersecting([0, 0], [8, 8], 1.0, [6, 6], [9, 9]) == ([6, 6], [7, 7])   # clipped


# ---- segment.test.cpp / triangle.test.cpp
def seg_clamp(seg, p):
    out = np.empty(2)
    ol.load().orc_segment_clamp(ol._dp(ol._arr(seg)), ol._dp(ol._arr(p)), ol._dp(out))
    return out


def tri_clamp(tri, p):
    out = np.empty(3)
    ol.load().orc_triangle_clamp(ol._dp(ol._arr(tri)), ol._dp(ol._arr(p)), ol._dp(out))
    return out


def intersects(face, c, radius):
    face = ol._arr(face)
    return bool(ol.load().orc_face_intersects(face.shape[1], ol._dp(face), ol._dp(ol._arr(c)), radius))


def winding(verts, faces, p):
    verts, faces = ol._arr(verts), ol._arr(faces, np.uint64)
    return ol.load().orc_winding(verts.shape[1], ol._dp(verts), len(verts), ol._u64p(faces), len(faces), ol._dp(ol._arr(p)))


def test_segment_clamp_and_intersects():
    seg = [(0.0, 0.0), (4.0, 4.0)]
    for p, q in (((0, 0), (0, 0)), ((-1, -1), (0, 0)), ((4, 4), (4, 4)), ((5, 5), (4, 4)), ((1, 1), (1, 1)), ((2, 2), (2, 2)),
                 ((1.9, 2.1), (2, 2)), ((2.1, 1.9), (2, 2))):
        assert approx(seg_clamp(seg, p), q), p
    deg = [(0.0, 0.0), (0.0, 0.0)]
    for p in ((0, 0), (1, 1), (-1, -1)):
        assert approx(seg_clamp(deg, p), (0, 0))
    face = [(0.0, 0.0), (2.0, 0.0)]
    assert intersects(face, (1.0, 0.0), 0.5)
    assert intersects(face, (0.0, 0.0), 0.5)
    assert intersects(face, (1.0, 0.5), 0.5)       # touching counts (inclusive)
    assert not intersects(face, (2.5, 0.5), 0.5)
    assert not intersects(face, (5.0, 5.0), 0.5)


def test_segment_and_triangle_winding_number():
    a, b, p = np.array([1.0, 0.0]), np.array([1.0, 1.0]), np.array([0.5, 0.5])
    v = lambda *pts: (np.array(pts), [list(range(len(pts)))])
    assert approx(winding(*v(a, b), p), 0.25)
    assert approx(winding(*v(b, a), p), -0.25)
    off = np.array([3.0, 4.0])
    assert approx(winding(*v(a + off, b + off), p + off), 0.25)
    assert approx(winding(*v(2 * a, 2 * b), 2 * p), 0.25)
    A, B, C, P = np.eye(3)[0], np.eye(3)[1], np.eye(3)[2], np.zeros(3)
    assert approx(winding(*v(A, B, C), P), 0.125)
    assert approx(winding(*v(C, B, A), P), -0.125)
    off = np.array([3.0, 4.0, 5.0])
    assert approx(winding(*v(A + off, B + off, C + off), P + off), 0.125)
    assert approx(winding(*v(2 * A, 2 * B, 2 * C), 2 * P), 0.125)


def test_triangle_clamp_and_intersects():
    tri = [(0.0, 0, 0), (2.0, 0, 0), (0, 2.0, 0)]
    table = [((-1, -1, 0), (0, 0, 0)), ((-0.5, -0.5, 5), (0, 0, 0)), ((3, -1, 0), (2, 0, 0)), ((2.5, -0.5, -3), (2, 0, 0)),
             ((-1, 3, 0), (0, 2, 0)), ((-0.5, 2.5, 2), (0, 2, 0)), ((1, -1, 0), (1, 0, 0)), ((0.5, -2, -1), (0.5, 0, 0)),
             ((-1, 1, 0), (0, 1, 0)), ((-2, 1.5, 4), (0, 1.5, 0)), ((2, 2, 0), (1, 1, 0)), ((1.5, 1.5, -2), (1, 1, 0)),
             ((0, 0, 0), (0, 0, 0)), ((2, 0, 0), (2, 0, 0)), ((0, 2, 0), (0, 2, 0)), ((0.5, 0.5, 0), (0.5, 0.5, 0)),
             ((1, 1, 1), (1, 1, 0)), ((1, 1, -1), (1, 1, 0))]
    for p, q in table:
        assert approx(tri_clamp(tri, p), q), p
    o, m, e = (0.0, 0, 0), (1.0, 1.0, 0), (2.0, 2.0, 0)
    for deg in ([o, o, e], [o, e, e], [o, e, o], [o, m, e]):
        for p, q in (((0, 0, 0), o), ((-1, -1, 0), o), ((-1, -1, -1), o), ((2, 2, 0), e), ((3, 3, 0), e), ((2, 2, 1), e),
                     ((1, 1, 0), m), ((0.9, 1.1, 0), m), ((1.1, 0.9, 0), m), ((1.1, 0.9, 1), m)):
            assert approx(tri_clamp(deg, p), q), (deg, p)
    for p in ((0, 0, 0), (1, 1, 1), (-1, -1, -1)):
        assert approx(tri_clamp([o, o, o], p), o)
    assert intersects(tri, (0, 0, 0), 0.5)
    assert intersects(tri, (0.5, 0.5, 0.2), 0.3)
    assert intersects(tri, (0.5, 0.5, 0.3), 0.3)
    assert not intersects(tri, (1.5, 1.5, 0), 0.1)
    assert not intersects(tri, (10, 10, 10), 0.5)
    assert not intersects(tri, (0.5, 0.5, 5), 0.5)


# ---- winding.test.cpp:62-100 (the exact winding number is the oracle's containment test)
def test_exact_winding_square_and_tetrahedron():
    sq = np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]])
    sf = [[0, 1], [1, 2], [2, 3], [3, 0]]
    assert approx(winding(sq, sf, (0.5, 0.5)), 1.0)
    assert approx(winding(sq, sf, (1.5, 0.5)), 0.0)
    assert approx(winding(sq, [f[::-1] for f in sf][::-1], (0.5, 0.5)), -1.0)
    tv = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    tf = [[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]]
    assert approx(winding(tv, tf, (0.25, 0.25, 0.25)), 1.0)
    assert approx(winding(tv, tf, (1.0, 1.0, 1.0)), 0.0)
    # degenerate faces contribute nothing (winding.test.cpp:69-79)
    dv = np.array([[0.0, 0, 0], [1, 0, 0], [2, 0, 0]])
    assert approx(winding(dv, [[0, 1, 2], [2, 1, 0]], (0.0, 1.0, 0.0)), 0.0)


# ---- face_search.test.cpp:108-160
def brute_faces(verts, faces, pts, radius):
    return [{f for f, idx in enumerate(faces) if intersects(verts[np.asarray(idx, int)], p, radius)} for p in pts]


@pytest.mark.parametrize("radius", [0.01, 0.1, 0.5, 1.0])
@pytest.mark.parametrize("scale", [0.5, 5.0])
def test_face_search_2d(radius, scale):
    pts = reference_cloud(2)
    verts, faces = cases.tessellate_2d(np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]]), np.array([[0, 1], [1, 2], [2, 3], [3, 0]], np.uint64), 0.05)
    assert len(faces) == 80 and len(verts) == 80      # 4 edges x ceil(1 / 0.05)
    off, cols = oracle_search(pts, radius, scale * radius, 2, verts, faces, want_faces=True)
    assert csr_rows(off, cols) == brute_faces(verts, faces, pts, radius)


@pytest.mark.parametrize("radius", [0.01, 0.1, 0.5, 1.0])
@pytest.mark.parametrize("scale", [0.5, 5.0])
def test_face_search_3d(radius, scale):
    """A structured box wall mesh and the reference's own set-up, a red-refined
    tetrahedron (face_search.test.cpp:136-160, cases.tessellate_3d), exercise the index."""
    pts = reference_cloud(3)
    verts, faces = cases._box_wall_mesh((1.0, 1.0, 1.0), (6, 5, 4))
    off, cols = oracle_search(pts, radius, scale * radius, 3, verts, faces, want_faces=True)
    assert csr_rows(off, cols) == brute_faces(verts, faces, pts, radius)
    tet = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    verts, faces = cases.tessellate_3d(tet, np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.uint64), 0.1)
    off, cols = oracle_search(pts, radius, scale * radius, 3, verts, faces, want_faces=True)
    assert csr_rows(off, cols) == brute_faces(verts, faces, pts, radius)
