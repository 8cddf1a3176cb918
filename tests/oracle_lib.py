"""ctypes binding of the CPU oracle (oracle/liboracle*.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")

FIELDS = {
    "m": 0, "gamma": 0, "rho": 0, "drho_dt": 0, "p": 0, "cs": 0, "phi": 0, "rho_raw": 0,
    "grad_gamma": 1, "grad_rho": 1, "v": 1, "dv_dt": 1, "r": 1, "dr": 1, "N": 1,
    "grad_v": 2, "L": 2,
}


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", _ODIR, "all"])


_libs = {}


def load(fast: bool = False) -> C.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_ODIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    d, vp, sz, u64p, dp = C.c_double, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.c_int] * 4
    lib.orc_destroy.argtypes = [vp]
    lib.orc_set_params.argtypes = [vp] + [d] * 8
    lib.orc_set_surface.argtypes = [vp, dp, sz, u64p, sz, dp, sz, u64p, sz]
    lib.orc_resize.argtypes = [vp, sz, sz]
    lib.orc_upload.argtypes = [vp, C.c_char_p, dp]
    lib.orc_download.argtypes = [vp, C.c_char_p, dp]
    lib.orc_stats.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.orc_phase_times.argtypes = [vp, dp, C.c_int]
    lib.orc_set_symmetric.argtypes = [vp, C.c_int]
    for f in ("orc_initialize", "orc_prepare", "orc_rhs_only", "orc_post_only"):
        getattr(lib, f).argtypes = [vp]
    lib.orc_step.argtypes = [vp, C.c_int, dp]
    lib.orc_neighbors.argtypes = [vp, u64p, u64p, sz, C.POINTER(sz)]
    lib.orc_face_neighbors.argtypes = [vp, u64p, u64p, sz, C.POINTER(sz)]
    lib.orc_kdtree_neighbors.argtypes = [C.c_int, dp, sz, d, u64p, u64p, sz, C.POINTER(sz)]
    lib.orc_set_num_threads.argtypes = [C.c_int]
    lib.orc_tiny.restype = d
    lib.orc_kernel_radius.restype = d
    lib.orc_kernel_radius.argtypes = [C.c_int, d]
    lib.orc_kernel_weight.restype = d
    lib.orc_kernel_weight.argtypes = [C.c_int, C.c_int]
    lib.orc_kernel_unit_value.restype = d
    lib.orc_kernel_unit_value.argtypes = [C.c_int, d]
    lib.orc_kernel_unit_deriv.restype = d
    lib.orc_kernel_unit_deriv.argtypes = [C.c_int, d]
    lib.orc_kernel_value.restype = d
    lib.orc_kernel_value.argtypes = [C.c_int, C.c_int, dp, d]
    lib.orc_kernel_grad.argtypes = [C.c_int, C.c_int, dp, d, dp]
    lib.orc_kernel_width_deriv.restype = d
    lib.orc_kernel_width_deriv.argtypes = [C.c_int, C.c_int, dp, d]
    lib.orc_kernel_antigrad.argtypes = [C.c_int, C.c_int, dp, d, dp]
    lib.orc_kernel_value_n.argtypes = [C.c_int, C.c_int, dp, sz, d, dp]
    lib.orc_kernel_antigrad_n.argtypes = [C.c_int, C.c_int, dp, sz, d, dp]
    lib.orc_kernel_flux.argtypes = [C.c_int, C.c_int, dp, dp, d, dp]
    lib.orc_kernel_antigrad_flux.restype = d
    lib.orc_kernel_antigrad_flux.argtypes = [C.c_int, C.c_int, dp, dp, d]
    lib.orc_segment_clamp.argtypes = [dp, dp, dp]
    lib.orc_triangle_clamp.argtypes = [dp, dp, dp]
    lib.orc_face_intersects.argtypes = [C.c_int, dp, dp, d]
    lib.orc_winding.restype = d
    lib.orc_winding.argtypes = [C.c_int, dp, sz, u64p, sz, dp]
    lib.orc_grid.argtypes = [C.c_int, dp, dp, d, u64p, dp]
    lib.orc_grid_cells_intersecting.argtypes = [C.c_int, dp, dp, d, dp, dp, u64p, u64p]
    lib.orc_lu_inverse.argtypes = [C.c_int, dp, dp]
    _libs[name] = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _arr(x, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


class OracleSolver:
    """Same Python surface as titsolver_b200.Solver, backed by the CPU oracle."""

    def __init__(self, dim, kernel_id=4, eos_id=0, integrator_id=3, fast=False):
        self.lib = load(fast)
        self.dim = dim
        self.h = self.lib.orc_create(dim, kernel_id, eos_id, integrator_id)
        assert self.h, "orc_create failed"
        self.n = 0

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, g, mu, cs0, rho0, xi, h, search_hint=0.0, face_hint=0.0):
        self.lib.orc_set_params(self.h, g, mu, cs0, rho0, xi, h, search_hint, face_hint)

    def set_surface(self, verts, faces, cverts, cfaces):
        v, f, cv, cf = _arr(verts), _arr(faces, np.uint64), _arr(cverts), _arr(cfaces, np.uint64)
        self.lib.orc_set_surface(self.h, _dp(v), len(v), _u64p(f), len(f), _dp(cv), len(cv), _u64p(cf), len(cf))

    def set_particles(self, n_fluid, n_fixed):
        self.n = n_fluid + n_fixed
        self.lib.orc_resize(self.h, n_fluid, n_fixed)

    def _shape(self, field):
        k = FIELDS[field]
        return (self.n,) if k == 0 else (self.n, self.dim) if k == 1 else (self.n, self.dim, self.dim)

    def upload(self, field, a):
        a = _arr(a)
        assert a.shape == self._shape(field), (field, a.shape, self._shape(field))
        assert self.lib.orc_upload(self.h, field.encode(), _dp(a)) == 0

    def download(self, field):
        a = np.empty(self._shape(field))
        assert self.lib.orc_download(self.h, field.encode(), _dp(a)) == 0
        return a

    def initialize(self):
        self.lib.orc_initialize(self.h)

    def prepare(self):
        self.lib.orc_prepare(self.h)

    def rhs_only(self):
        self.lib.orc_rhs_only(self.h)

    def post_only(self):
        self.lib.orc_post_only(self.h)

    STAT_NAMES = ("lu_failed", "free_surface", "splash", "near_surface", "shifted", "shifted_no_v_correction", "fs_corrected", "fs_candidates")

    def stats(self):
        """Branch counters of the last post_integrate (oracle_sim.h, SimBase::stats)."""
        a = (C.c_longlong * 8)()
        self.lib.orc_stats(self.h, a)
        return dict(zip(self.STAT_NAMES, [int(x) for x in a]))

    PHASE_NAMES = ("search", "compute_gamma", "setup_boundary", "continuity_momentum", "update_lincomb_dt", "apply_shifts", "free_surface_correction")

    def set_symmetric(self, on=True):
        """Pair sums over unordered pairs, both particles updated per pair, block-coloured (the reference's
        loop structure; CPU baseline of bench.py). Default off: gather form with a fixed order of every sum."""
        self.lib.orc_set_symmetric(self.h, int(bool(on)))

    def phase_times(self, reset=True):
        """Seconds per phase accumulated since the last reset (oracle_sim.h, SimBase::phase_s)."""
        a = (C.c_double * 8)()
        self.lib.orc_phase_times(self.h, a, int(reset))
        return dict(zip(self.PHASE_NAMES, [float(x) for x in a]))

    def step(self, nsteps=1):
        dt = C.c_double(0)
        self.lib.orc_step(self.h, nsteps, C.byref(dt))
        return dt.value

    def _csr(self, fn):
        nnz = C.c_size_t(0)
        fn(self.h, None, None, 0, C.byref(nnz))
        off = np.zeros(self.n + 1, np.uint64)
        cols = np.zeros(max(nnz.value, 1), np.uint64)
        assert fn(self.h, _u64p(off), _u64p(cols), nnz.value, C.byref(nnz)) == 0
        return off, cols[: nnz.value]

    def neighbors(self):
        return self._csr(self.lib.orc_neighbors)

    def face_neighbors(self):
        return self._csr(self.lib.orc_face_neighbors)


def kdtree_neighbors(pts, radius):
    """Sorted neighbour rows (CSR) through the oracle's K-d tree index
    (geom/search/kd_tree_search.hpp restated in oracle_math.h)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n, dim = pts.shape
    lib = load()
    nnz = C.c_size_t(0)
    assert lib.orc_kdtree_neighbors(dim, _dp(pts), n, radius, None, None, 0, C.byref(nnz)) == 0
    off = np.zeros(n + 1, np.uint64)
    cols = np.zeros(max(nnz.value, 1), np.uint64)
    assert lib.orc_kdtree_neighbors(dim, _dp(pts), n, radius, _u64p(off), _u64p(cols), nnz.value, C.byref(nnz)) == 0
    return off, cols[: nnz.value]


def load_case(solver, case):
    """Feed a titsolver_b200.cases.Case into a solver (oracle or GPU)."""
    solver.set_params(case.g, case.mu, case.cs0, case.rho0, case.xi, case.h)
    solver.set_surface(case.verts, case.faces, case.cverts, case.cfaces)
    solver.set_particles(case.n_fluid, case.n_fixed)
    solver.upload("r", case.r)
    solver.upload("m", case.m)
    solver.upload("rho", case.rho)


# ---- kernel-layer helpers -----------------------------------------------------
def kvalue(kid, x, h):
    x = _arr(x)
    return load().orc_kernel_value(kid, len(x), _dp(x), h)


def kgrad(kid, x, h):
    x = _arr(x)
    out = np.empty(len(x))
    load().orc_kernel_grad(kid, len(x), _dp(x), h, _dp(out))
    return out


def kwidth_deriv(kid, x, h):
    x = _arr(x)
    return load().orc_kernel_width_deriv(kid, len(x), _dp(x), h)


def kantigrad(kid, x, h):
    x = _arr(x)
    out = np.empty(len(x))
    load().orc_kernel_antigrad(kid, len(x), _dp(x), h, _dp(out))
    return out


def kflux(kid, face, x, h):
    face, x = _arr(face), _arr(x)
    out = np.empty(len(x))
    load().orc_kernel_flux(kid, len(x), _dp(face), _dp(x), h, _dp(out))
    return out


def kantigrad_flux(kid, face, x, h):
    face, x = _arr(face), _arr(x)
    return load().orc_kernel_antigrad_flux(kid, len(x), _dp(face), _dp(x), h)
