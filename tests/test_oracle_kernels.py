"""Pins the oracle's kernel layer against the reference's own known-answer tests.

Restates /root/reference/source/tit/sph/kernel.test.cpp (all six kernels):
  :33-91    normalisation over sphere (2D, 3D) and box (1D, 2D, 3D), h in {1, .1, .01}
  :95-122   grad vs the exact derivative of the value
  :126-139  width_deriv vs the exact derivative in h
  :143-180  antigrad: divergence identity div A = W
  :184-300  flux vs quadrature for segments / triangles + flip antisymmetry
  :304-542  antigrad_flux vs quadrature + closed forms 0, 1/4, 1/2 (2D), 1/12, 1/4, 1/2 (3D)
with the reference's tolerance: absolute tiny = cbrt(eps) (testing/test.hpp:42,
core/math.hpp:145-162).

The reference differentiates with dual numbers and integrates with an adaptive
degree-5 cubature (testing/math/integrals.hpp:129-224). Here the "exact" side
is an INDEPENDENT mpmath restatement of the kernel definitions
(sph/kernel.gen.cpp:520-554) at 40 digits, and the quadrature is an adaptive
tensor Gauss-Legendre cubature with the same accept/split rule.
"""
import itertools

import mpmath as mp
import numpy as np
import pytest

import oracle_lib as ol

mp.mp.dps = 40
TINY = float(np.cbrt(np.finfo(float).eps))
KERNELS = list(range(6))
HS = (1.0, 0.1, 0.01)

# (cutoff, piece) lists exactly as kernel.gen.cpp:520-554 defines them.
Q = mp.mpf
PIECES = {
    0: [(2, lambda q: Q(1) / 4 * (2 - q) ** 3), (1, lambda q: -((1 - q) ** 3))],
    1: [(Q(5) / 2, lambda q: (Q(5) / 2 - q) ** 4), (Q(3) / 2, lambda q: -5 * (Q(3) / 2 - q) ** 4), (Q(1) / 2, lambda q: 10 * (Q(1) / 2 - q) ** 4)],
    2: [(3, lambda q: (3 - q) ** 5), (2, lambda q: -6 * (2 - q) ** 5), (1, lambda q: 15 * (1 - q) ** 5)],
    3: [(2, lambda q: (1 + 2 * q) * (1 - q / 2) ** 4)],
    4: [(2, lambda q: (1 + 3 * q + Q(35) / 12 * q**2) * (1 - q / 2) ** 6)],
    5: [(2, lambda q: (1 + 4 * q + Q(25) / 4 * q**2 + 4 * q**3) * (1 - q / 2) ** 8)],
}
UNIT_RADIUS = {0: 2.0, 1: 2.5, 2: 3.0, 3: 2.0, 4: 2.0, 5: 2.0}
SPHERE_AREA = {1: Q(2), 2: 2 * mp.pi, 3: 4 * mp.pi}


def w_exact(kid, q):
    return sum((f(q) for c, f in PIECES[kid] if q < c), Q(0))


def breaks(kid, lo=0):
    return sorted({Q(lo)} | {Q(c) for c, _ in PIECES[kid] if c > lo})


def moment_exact(kid, dim, q):
    """m_D(q) = int_q^R xi^(D-1) w(xi) d xi (sph/kernel.hpp:127-131)."""
    q = Q(q)
    pts = [q] + [b for b in breaks(kid) if b > q]
    if len(pts) < 2:
        return Q(0)
    return mp.quad(lambda t: t ** (dim - 1) * w_exact(kid, t), pts)


_weights = {}


def weight_exact(kid, dim):
    if (kid, dim) not in _weights:
        _weights[kid, dim] = 1 / (moment_exact(kid, dim, 0) * SPHERE_AREA[dim])
    return _weights[kid, dim]


def W_exact(kid, x, h):
    dim = len(x)
    q = mp.sqrt(sum(Q(c) ** 2 for c in x)) / Q(h)
    return weight_exact(kid, dim) / Q(h) ** dim * w_exact(kid, q)


def antigrad_exact(kid, x, h):
    dim = len(x)
    q = mp.sqrt(sum(Q(c) ** 2 for c in x)) / Q(h)
    c = -weight_exact(kid, dim) / Q(h) ** dim * moment_exact(kid, dim, q) / q**dim
    return [Q(xi) * c for xi in x]


# ---- adaptive cubature (same accept/split rule as testing/math/integrals.hpp:129-146)
_GX, _GW = np.polynomial.legendre.leggauss(5)
_GX, _GW = 0.5 * (_GX + 1.0), 0.5 * _GW


def _panel_sums(f, lo, size):
    """Tensor Gauss on each panel [lo, lo + size]; f maps (m, dim) points -> (m,) values."""
    npan, dim = lo.shape
    grids = np.meshgrid(*([_GX] * dim), indexing="ij")
    ref = np.stack([g.ravel() for g in grids], axis=1)                       # (G^dim, dim)
    wts = np.prod(np.stack(np.meshgrid(*([_GW] * dim), indexing="ij")), axis=0).ravel()
    pts = lo[:, None, :] + ref[None, :, :] * size[:, None, :]
    vals = f(pts.reshape(-1, dim)).reshape(npan, -1)
    return vals @ wts * np.prod(size, axis=1)


def integrate_box(f, lo, hi, eps=TINY, max_depth=14):
    lo, hi = np.atleast_1d(np.asarray(lo, float)), np.atleast_1d(np.asarray(hi, float))
    dim = len(lo)
    corners = np.array(list(itertools.product((0.0, 0.5), repeat=dim)))
    plo, psz = lo[None, :], (hi - lo)[None, :]
    est = _panel_sums(f, plo, psz)
    tol = np.array([eps])
    total = 0.0
    for _ in range(max_depth):
        if len(plo) == 0:
            return total
        clo = (plo[:, None, :] + corners[None, :, :] * psz[:, None, :]).reshape(-1, dim)
        csz = np.repeat(psz * 0.5, len(corners), axis=0)
        cest = _panel_sums(f, clo, csz)
        fine = cest.reshape(len(plo), -1).sum(axis=1)
        ok = np.abs(fine - est) <= tol
        total += fine[ok].sum()
        keep = np.repeat(~ok, len(corners))
        plo, psz, est = clo[keep], csz[keep], cest[keep]
        tol = np.repeat(tol[~ok] / 2**dim, len(corners))
    return total + est.sum()


def value_n(kid, pts, h):
    pts = np.ascontiguousarray(pts)
    out = np.empty(len(pts))
    ol.load().orc_kernel_value_n(kid, pts.shape[1], ol._dp(pts), len(pts), h, ol._dp(out))
    return out


def antigrad_n(kid, pts, h):
    pts = np.ascontiguousarray(pts)
    out = np.empty_like(pts)
    ol.load().orc_kernel_antigrad_n(kid, pts.shape[1], ol._dp(pts), len(pts), h, ol._dp(out))
    return out


def approx(a, b, eps=TINY):
    """approx_equal_to: |a - b| <= tiny for scalars (core/math.hpp:156-162), Euclidean
    norm for vectors (core/_vec/vec.hpp:681-687)."""
    d = np.atleast_1d(np.asarray(a, float) - np.asarray(b, float))
    return float(np.dot(d, d)) <= eps * eps


# ---- the oracle's constants and radial functions against the exact definitions
@pytest.mark.parametrize("kid", KERNELS)
def test_weights_radius_and_radial_functions(kid):
    lib = ol.load()
    assert lib.orc_kernel_radius(kid, 0.37) == pytest.approx(UNIT_RADIUS[kid] * 0.37, rel=1e-15)
    for dim in (1, 2, 3):
        assert lib.orc_kernel_weight(kid, dim) == pytest.approx(float(weight_exact(kid, dim)), rel=1e-14)
    for q in np.linspace(0.0, UNIT_RADIUS[kid] + 0.25, 67):
        we = w_exact(kid, Q(float(q)))
        de = mp.diff(lambda t: w_exact(kid, t), Q(float(q))) if all(abs(q - float(c)) > 1e-9 for c, _ in PIECES[kid]) else None
        assert abs(lib.orc_kernel_unit_value(kid, q) - float(we)) <= 1e-13 * max(1.0, abs(float(we)))
        if de is not None and q > 0:
            assert abs(lib.orc_kernel_unit_deriv(kid, q) - float(de)) <= 1e-12 * max(1.0, abs(float(de)))


def test_wendland_c4_constants_match_survey():
    """omega_2 = 9/(4 pi), omega_3 = 495/(256 pi)."""
    assert float(weight_exact(4, 2)) == pytest.approx(9 / (4 * np.pi), rel=1e-15)
    assert float(weight_exact(4, 3)) == pytest.approx(495 / (256 * np.pi), rel=1e-15)
    assert float(weight_exact(0, 2)) == pytest.approx(10 / (7 * np.pi), rel=1e-15)


# ---- kernel.test.cpp:33-91
@pytest.mark.parametrize("kid", KERNELS)
@pytest.mark.parametrize("h", HS)
def test_normalised_over_sphere(kid, h):
    r = ol.load().orc_kernel_radius(kid, h)

    def polar(p):
        x = np.stack([p[:, 0] * np.cos(p[:, 1]), p[:, 0] * np.sin(p[:, 1])], axis=1)
        return p[:, 0] * value_n(kid, x, h)

    assert approx(integrate_box(polar, [0, 0], [r, 2 * np.pi]), 1.0)

    def spherical(p):
        rr, th, ph = p[:, 0], p[:, 1], p[:, 2]
        x = np.stack([rr * np.sin(th) * np.cos(ph), rr * np.sin(th) * np.sin(ph), rr * np.cos(th)], axis=1)
        return rr**2 * np.sin(th) * value_n(kid, x, h)

    assert approx(integrate_box(spherical, [0, 0, 0], [r, np.pi, 2 * np.pi]), 1.0)


@pytest.mark.parametrize("kid", KERNELS)
@pytest.mark.parametrize("h", HS)
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_normalised_over_box(kid, h, dim):
    r = ol.load().orc_kernel_radius(kid, h)
    # slightly larger than the support: also checks that the kernel vanishes outside
    assert approx(integrate_box(lambda p: value_n(kid, p, h), [-r] * dim, [r] * dim), 1.0)


# ---- kernel.test.cpp:95-139
@pytest.mark.parametrize("kid", KERNELS)
@pytest.mark.parametrize("h", HS)
def test_grad_and_width_deriv(kid, h):
    assert ol.load().orc_kernel_radius(kid, h) >= h * np.sqrt(3.0)
    x = [h * h * 0.1] * 3
    g = ol.kgrad(kid, x, h)
    for i in range(3):
        def along(t, i=i):
            y = [Q(c) for c in x]
            y[i] = t
            return W_exact(kid, y, h)
        d = float(mp.diff(along, Q(x[i])))
        assert approx(g[i], d)
        assert abs(g[i] - d) <= 1e-11 * abs(d) + 1e-300
    dh = float(mp.diff(lambda t: W_exact(kid, x, t), Q(h)))
    got = ol.kwidth_deriv(kid, x, h)
    assert approx(got, dh)
    assert abs(got - dh) <= 1e-11 * abs(dh)
    assert ol.kvalue(kid, x, h) == pytest.approx(float(W_exact(kid, x, h)), rel=1e-13)


# ---- kernel.test.cpp:143-180
@pytest.mark.parametrize("kid", KERNELS)
@pytest.mark.parametrize("h", HS)
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_antigrad_divergence_identity(kid, h, dim):
    x = [h * c for c in (0.37, 0.23, 0.11)[:dim]]
    # (1) the oracle's antigradient equals the exact formula (kernel.hpp:182-191) ...
    a, e = ol.kantigrad(kid, x, h), antigrad_exact(kid, x, h)
    for i in range(dim):
        assert abs(a[i] - float(e[i])) <= 1e-12 * abs(float(e[i]))
    # (2) ... whose divergence is the kernel value — the reference's check.
    div = Q(0)
    for i in range(dim):
        def comp(t, i=i):
            y = [Q(c) for c in x]
            y[i] = t
            return antigrad_exact(kid, y, h)[i]
        div += mp.diff(comp, Q(x[i]))
    assert approx(float(div), ol.kvalue(kid, x, h))
    assert abs(float(div) - ol.kvalue(kid, x, h)) <= 1e-10 * abs(float(div))


# ---- face geometry helpers (geom/segment.hpp:61-68, geom/triangle.hpp:78-100)
def normalize(v):
    """core/_vec/vec.hpp:663-678: vectors shorter than tiny normalise to ZERO. The
    reference's flux tests rely on it: the h/6 triangle at h = 0.01 has
    |wnormal| = 2.4e-6 < tiny, so both flux() and the estimate are exactly 0."""
    n2 = float(np.dot(v, v))
    return v / np.sqrt(n2) if n2 >= TINY * TINY else np.zeros_like(v)


def seg_normal(a, b):
    ba = np.asarray(b, float) - np.asarray(a, float)
    return normalize(np.array([ba[1], -ba[0]]))


def tri_normal(a, b, c):
    return normalize(np.cross(np.asarray(b, float) - a, np.asarray(c, float) - a) / 2)


def flux_estimate_2d(kid, seg, x, h):
    a, b = np.asarray(seg[0], float), np.asarray(seg[1], float)
    L = np.linalg.norm(b - a)
    f = lambda t: L * value_n(kid, np.asarray(x)[None, :] - (a[None, :] + t * (b - a)[None, :]), h)
    return seg_normal(a, b) * integrate_box(f, [0], [1])


def antiflux_estimate_2d(kid, seg, x, h):
    a, b = np.asarray(seg[0], float), np.asarray(seg[1], float)
    L, n = np.linalg.norm(b - a), seg_normal(a, b)
    f = lambda t: L * (antigrad_n(kid, np.asarray(x)[None, :] - (a[None, :] + t * (b - a)[None, :]), h) @ n)
    return -integrate_box(f, [0], [1])


def _tri_points(tri, t):
    a, b, c = (np.asarray(v, float) for v in tri)
    return a[None, :] + t[:, :1] * ((1 - t[:, 1:2]) * (b - a)[None, :] + t[:, 1:2] * (c - a)[None, :])


def tri_area(tri):
    a, b, c = (np.asarray(v, float) for v in tri)
    return 0.5 * np.linalg.norm(np.cross(b - a, c - a))


def flux_estimate_3d(kid, tri, x, h):
    J = tri_area(tri)
    f = lambda t: 2 * t[:, 0] * J * value_n(kid, np.asarray(x)[None, :] - _tri_points(tri, t), h)
    return tri_normal(*tri) * integrate_box(f, [0, 0], [1, 1])


def antiflux_estimate_3d(kid, tri, x, h):
    J, n = tri_area(tri), tri_normal(*tri)
    f = lambda t: 2 * t[:, 0] * J * (antigrad_n(kid, np.asarray(x)[None, :] - _tri_points(tri, t), h) @ n)
    return -integrate_box(f, [0, 0], [1, 1])


OS = (1.0, 0.9, 0.5, 0.1, 0.0, -0.1, -0.5, -0.9, -1.0)
SS = (1.0, 0.9, 0.5, 0.1, 0.0)


# ---- kernel.test.cpp:184-236
@pytest.mark.parametrize("kid", KERNELS)
def test_flux_2d(kid):
    for h in HS:
        seg = [(0.0, 0.0), (2.0, 0.0)]
        assert approx(flux_estimate_2d(kid, seg, (10.0, 10.0), h), 0.0)
        assert approx(ol.kflux(kid, seg, (10.0, 10.0), h), 0.0)
        for seg, xf in (([(0.0, 0.0), (2.0 * h, 0.0)], lambda o, s: (1.0 + o * h, s * h)),
                        ([(-h / 4, 0.0), (h / 4, 0.0)], lambda o, s: (o * h, s * h))):
            flipped = seg[::-1]
            for o in OS:
                for s in SS:
                    x = xf(o, s)
                    got = ol.kflux(kid, seg, x, h)
                    assert approx(got, flux_estimate_2d(kid, seg, x, h)), (h, o, s)
                    assert approx(got, -ol.kflux(kid, flipped, x, h)), (h, o, s)


# ---- kernel.test.cpp:237-300
@pytest.mark.parametrize("kid", KERNELS)
def test_flux_3d(kid):
    for h in HS:
        tri = [(2.0, 0, 0), (0, 2.0, 0), (0, 0, 2.0)]
        assert approx(flux_estimate_3d(kid, tri, (10.0, 10.0, 10.0), h), 0.0)
        assert approx(ol.kflux(kid, tri, (10.0, 10.0, 10.0), h), 0.0)
        for tri, xf in (([(3 * h, 0, 0), (0, 3 * h, 0), (0, 0, 3 * h)], lambda o, s: (1.0 + o * h, 1.0 + s * h, 1.0 - s * h)),
                        ([(h / 6, 0, 0), (0, h / 6, 0), (0, 0, h / 6)], lambda o, s: (o * h, s * h, s * h))):
            flipped = tri[::-1]
            for o in OS:
                for s in SS:
                    x = xf(o, s)
                    got = ol.kflux(kid, tri, x, h)
                    assert approx(got, flux_estimate_3d(kid, tri, x, h)), (h, o, s)
                    assert approx(got, -ol.kflux(kid, flipped, x, h)), (h, o, s)


# ---- kernel.test.cpp:304-394
@pytest.mark.parametrize("kid", KERNELS)
def test_antigrad_flux_2d(kid):
    for h in HS:
        seg = [(0.0, 0.0), (2.0, 0.0)]
        assert approx(ol.kantigrad_flux(kid, seg, (10.0, 10.0), h), 0.0)
        for seg, xf in (([(0.0, 0.0), (2.0 * h, 0.0)], lambda o, s: (1.0 + o * h, s * h)),
                        ([(-h / 2, 0.0), (h / 2, 0.0)], lambda o, s: (o * h, s * h))):
            flipped = seg[::-1]
            for o in OS:
                for s in SS[:4]:
                    x = xf(o, s)
                    got = ol.kantigrad_flux(kid, seg, x, h)
                    assert approx(got, antiflux_estimate_2d(kid, seg, x, h)), (h, o, s)
                    assert approx(got, -ol.kantigrad_flux(kid, flipped, x, h)), (h, o, s)
    # point on the segment line: closed forms (kernel.test.cpp:357-393)
    a, b = np.array([0.0, 0.0]), np.array([2.0, 2.0])
    seg = [tuple(a), tuple(b)]
    for h in HS:
        for o in (1.0, 0.9, 0.5, 0.1):
            assert approx(ol.kantigrad_flux(kid, seg, a - o, h), 0.0)
            assert approx(ol.kantigrad_flux(kid, seg, b + o, h), 0.0)
        assert approx(abs(ol.kantigrad_flux(kid, seg, a, h)), 0.25)
        assert approx(abs(ol.kantigrad_flux(kid, seg, b, h)), 0.25)
        for o in (0.9, 0.5, 0.1, 0.0, -0.1, -0.5, -0.9):
            t = 0.5 * (o + 1.0)
            assert approx(abs(ol.kantigrad_flux(kid, seg, a + t * (b - a), h)), 0.5)


# ---- kernel.test.cpp:395-542
@pytest.mark.parametrize("kid", KERNELS)
def test_antigrad_flux_3d(kid):
    O3 = (1.0, 0.9, 0.5, 0.1, -0.1, -0.5, -0.9, -1.0)
    for h in HS:
        tri = [(2.0, 0, 0), (0, 2.0, 0), (0, 0, 2.0)]
        assert approx(ol.kantigrad_flux(kid, tri, (10.0, 10.0, 10.0), h), 0.0)
        for tri, xf in (([(3 * h, 0, 0), (0, 3 * h, 0), (0, 0, 3 * h)], lambda o, s: (1.0 + o * h, 1.0 + s * h, 1.0 - s * h)),
                        ([(h / 6, 0, 0), (0, h / 6, 0), (0, 0, h / 6)], lambda o, s: (o * h, s * h, -s * h))):
            flipped = tri[::-1]
            for o in O3:
                for s in SS[:4]:
                    x = xf(o, s)
                    got = ol.kantigrad_flux(kid, tri, x, h)
                    assert approx(got, antiflux_estimate_3d(kid, tri, x, h)), (h, o, s)
                    assert approx(got, -ol.kantigrad_flux(kid, flipped, x, h)), (h, o, s)
    # point on the triangle plane: closed forms (kernel.test.cpp:476-539)
    A, B, Cc = np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([0, 0, 1.0])
    tri = [tuple(A), tuple(B), tuple(Cc)]
    ctr = (A + B + Cc) / 3
    for h in HS:
        for o in (1.0, 0.9, 0.5, 0.1):
            assert approx(ol.kantigrad_flux(kid, tri, A + np.array([2 * o, -o, -o]), h), 0.0)
            assert approx(ol.kantigrad_flux(kid, tri, B + np.array([-o, 2 * o, -o]), h), 0.0)
            assert approx(ol.kantigrad_flux(kid, tri, Cc + np.array([-o, -o, 2 * o]), h), 0.0)
        for v in (A, B, Cc):
            assert approx(ol.kantigrad_flux(kid, tri, v, h), 1.0 / 12.0)
        for o in (0.9, 0.5, 0.1, 0.0, -0.1, -0.5, -0.9):
            t = 0.5 * (o + 1.0)
            assert approx(abs(ol.kantigrad_flux(kid, tri, A + t * (B - A), h)), 0.25)
            assert approx(abs(ol.kantigrad_flux(kid, tri, A + t * (Cc - A), h)), 0.25)
            assert approx(abs(ol.kantigrad_flux(kid, tri, B + t * (Cc - B), h)), 0.25)
            for v in (A, B, Cc):
                assert approx(abs(ol.kantigrad_flux(kid, tri, v + t * (ctr - v), h)), 0.5)
