"""titsolver_b200 — B200-native WCSPH particle step behind the TitSolver API.

The product is `libtitgpu.so` (hand-written sm_100a CUDA kernels + a C ABI,
see include/titgpu.h) and the C++ facade in include/tit_b200/. This module is the
thin ctypes binding used by the tests and bench.py; it mirrors the call
sequence of /root/reference/source/titwcsph/wcsph.cpp.

There is no CPU fallback: importing works without a GPU (so that the ABI can be
inspected), every compute call requires one and raises `TitGpuError` otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import cases  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# TITGPU_LIB selects a tuning variant of the SAME CUDA library (see build.py); there is no other backend.
LIB_PATH = os.environ.get("TITGPU_LIB") or os.path.join(_HERE, "libtitgpu.so")

FIELDS = {
    "m": 0, "gamma": 0, "rho": 0, "drho_dt": 0, "p": 0, "cs": 0, "phi": 0, "rho_raw": 0,
    "grad_gamma": 1, "grad_rho": 1, "v": 1, "dv_dt": 1, "r": 1, "dr": 1, "N": 1,
    "grad_v": 2, "L": 2,
}

ABI_SYMBOLS = [
    "titgpu_create", "titgpu_destroy", "titgpu_last_error", "titgpu_set_params", "titgpu_set_surface",
    "titgpu_upload", "titgpu_download", "titgpu_initialize", "titgpu_prepare", "titgpu_rhs_only", "titgpu_step",
    "titgpu_set_outputs", "titgpu_set_lists", "titgpu_list_redos", "titgpu_set_tiles", "titgpu_set_group_sweep", "titgpu_set_graphs", "titgpu_graph_replays", "titgpu_mg_reserve", "titgpu_mg_counts",
    "titgpu_mg_set_slab", "titgpu_mg_set_halo_pair", "titgpu_mg_set_gids", "titgpu_mg_attach_comm", "titgpu_mg_nccl_unique_id", "titgpu_mg_attach_nccl", "titgpu_mg_hub_create", "titgpu_mg_hub_destroy",
    "titgpu_mg_attach_hub", "titgpu_mg_detach", "titgpu_mg_download_owned", "titgpu_mg_upload_owned", "titgpu_mg_stats",
    "titgpu_neighbors", "titgpu_face_neighbors", "titgpu_synchronize", "titgpu_launch_count", "titgpu_stream", "titgpu_version",
    "titgpu_profile_enable", "titgpu_profile_reset", "titgpu_profile_count", "titgpu_profile_get", "titgpu_measure_fp64_peak",
]


class TitGpuError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libtitgpu.so; fails loudly when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TitGpuError(f"{LIB_PATH} is missing: build it with `python -m titsolver_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    d, vp, sz = C.c_double, C.c_void_p, C.c_size_t
    u64p = C.POINTER(C.c_uint64)
    lib.titgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.titgpu_destroy.argtypes = [vp]
    lib.titgpu_last_error.argtypes = [vp]
    lib.titgpu_last_error.restype = C.c_char_p
    lib.titgpu_set_params.argtypes = [vp] + [d] * 8
    lib.titgpu_set_surface.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz]
    lib.titgpu_upload.argtypes = [vp, sz, sz, C.c_char_p, vp, sz]
    lib.titgpu_download.argtypes = [vp, C.c_char_p, vp, sz]
    for f in ("titgpu_initialize", "titgpu_prepare", "titgpu_rhs_only", "titgpu_synchronize"):
        getattr(lib, f).argtypes = [vp]
    lib.titgpu_step.argtypes = [vp, C.c_int, C.POINTER(d)]
    lib.titgpu_neighbors.argtypes = [vp, u64p, u64p, sz, C.POINTER(sz)]
    lib.titgpu_face_neighbors.argtypes = [vp, u64p, u64p, sz, C.POINTER(sz)]
    lib.titgpu_set_outputs.argtypes = [vp, C.c_int]
    lib.titgpu_set_lists.argtypes = [vp, C.c_int]
    lib.titgpu_set_tiles.argtypes = [vp, C.c_int]
    lib.titgpu_set_group_sweep.argtypes = [vp, C.c_int]
    lib.titgpu_set_graphs.argtypes = [vp, C.c_int]
    lib.titgpu_graph_replays.argtypes = [vp]
    lib.titgpu_graph_replays.restype = C.c_ulonglong
    lib.titgpu_list_redos.argtypes = [vp]
    lib.titgpu_list_redos.restype = C.c_ulonglong
    lib.titgpu_mg_reserve.argtypes = [vp, sz]
    lib.titgpu_mg_counts.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    lib.titgpu_mg_set_slab.argtypes = [vp, C.c_int, d, d, d, C.c_longlong]
    lib.titgpu_mg_set_halo_pair.argtypes = [vp, d]
    lib.titgpu_mg_set_gids.argtypes = [vp, vp]
    lib.titgpu_mg_attach_comm.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.titgpu_mg_nccl_unique_id.argtypes = [vp]
    lib.titgpu_mg_attach_nccl.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.titgpu_mg_hub_create.argtypes = [C.c_int]
    lib.titgpu_mg_hub_create.restype = vp
    lib.titgpu_mg_hub_destroy.argtypes = [vp]
    lib.titgpu_mg_hub_destroy.restype = None
    lib.titgpu_mg_attach_hub.argtypes = [vp, vp, C.c_int]
    lib.titgpu_mg_detach.argtypes = [vp]
    lib.titgpu_mg_download_owned.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.titgpu_mg_upload_owned.argtypes = [vp, sz, vp, vp, vp]
    lib.titgpu_mg_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.titgpu_launch_count.argtypes = [vp]
    lib.titgpu_launch_count.restype = C.c_ulonglong
    lib.titgpu_stream.argtypes = [vp]
    lib.titgpu_stream.restype = vp
    lib.titgpu_version.restype = C.c_char_p
    lib.titgpu_profile_enable.argtypes = [vp, C.c_int]
    lib.titgpu_profile_reset.argtypes = [vp]
    lib.titgpu_profile_count.argtypes = [vp]
    lib.titgpu_profile_get.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_ulonglong), C.POINTER(d)]
    lib.titgpu_measure_fp64_peak.argtypes = [vp, C.POINTER(d)]
    _lib = lib
    return lib


def nccl_unique_id() -> bytes:
    """An ncclUniqueId (128 bytes) for titgpu_mg_attach_nccl; make it on one rank, send it to the others."""
    buf = C.create_string_buffer(128)
    if load_library().titgpu_mg_nccl_unique_id(buf):
        raise TitGpuError("titgpu_mg_nccl_unique_id failed (is libnccl.so.2 loadable?)")
    return buf.raw


def hub_create(nranks):
    h = load_library().titgpu_mg_hub_create(int(nranks))
    if not h:
        raise TitGpuError("titgpu_mg_hub_create failed")
    return h


def hub_destroy(hub):
    load_library().titgpu_mg_hub_destroy(hub)


def _arr(x, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


class Solver:
    """One GPU context == FluidEquations + integrator + ParticleArray + ParticleMesh
    of the reference driver (wcsph.cpp:79-154)."""

    def __init__(self, dim, kernel_id=4, eos_id=0, integrator_id=3, device=0):
        self.lib = load_library()
        self.dim = dim
        self.h = C.c_void_p()
        rc = self.lib.titgpu_create(C.byref(self.h), device, dim, kernel_id, eos_id, integrator_id)
        if rc:
            msg = self._err()
            if self.h:
                self.lib.titgpu_destroy(self.h)
                self.h = C.c_void_p()
            raise TitGpuError(f"titgpu_create failed: {msg}")
        self.n_fluid = self.n_fixed = 0

    @property
    def n(self):
        return self.n_fluid + self.n_fixed

    def _err(self):
        return (self.lib.titgpu_last_error(self.h) or b"").decode() if self.h else "no context"

    def _ck(self, rc, what):
        if rc:
            raise TitGpuError(f"{what} failed: {self._err()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.titgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, g, mu, cs0, rho0, xi, h, search_hint=0.0, face_hint=0.0):
        self._ck(self.lib.titgpu_set_params(self.h, g, mu, cs0, rho0, xi, h, search_hint, face_hint), "titgpu_set_params")

    def set_surface(self, verts, faces, cverts, cfaces):
        v, f, cv, cf = _arr(verts), _arr(faces, np.uint64), _arr(cverts), _arr(cfaces, np.uint64)
        self._ck(self.lib.titgpu_set_surface(self.h, v.ctypes.data, len(v), f.ctypes.data, len(f), cv.ctypes.data, len(cv), cf.ctypes.data, len(cf)), "titgpu_set_surface")

    def set_particles(self, n_fluid, n_fixed):
        self.n_fluid, self.n_fixed = int(n_fluid), int(n_fixed)

    def _shape(self, field):
        k = FIELDS[field]
        return (self.n,) if k == 0 else (self.n, self.dim) if k == 1 else (self.n, self.dim, self.dim)

    def upload(self, field, a, stride_bytes=0):
        a = _arr(a)
        if stride_bytes == 0:
            assert a.shape == self._shape(field), (field, a.shape, self._shape(field))
        self._ck(self.lib.titgpu_upload(self.h, self.n_fluid, self.n_fixed, field.encode(), a.ctypes.data, stride_bytes), f"titgpu_upload({field})")

    def upload_raw(self, field, ptr, stride_bytes=0):
        """Upload from a raw host address (e.g. pinned memory)."""
        self._ck(self.lib.titgpu_upload(self.h, self.n_fluid, self.n_fixed, field.encode(), ptr, stride_bytes), f"titgpu_upload({field})")

    def download(self, field, out=None):
        if out is None:
            out = np.empty(self._shape(field))
        self._ck(self.lib.titgpu_download(self.h, field.encode(), out.ctypes.data, 0), f"titgpu_download({field})")
        return out

    def download_raw(self, field, ptr, stride_bytes=0):
        self._ck(self.lib.titgpu_download(self.h, field.encode(), ptr, stride_bytes), f"titgpu_download({field})")

    def initialize(self):
        self._ck(self.lib.titgpu_initialize(self.h), "titgpu_initialize")

    def prepare(self):
        self._ck(self.lib.titgpu_prepare(self.h), "titgpu_prepare")

    def rhs_only(self):
        self._ck(self.lib.titgpu_rhs_only(self.h), "titgpu_rhs_only")

    def step(self, nsteps=1):
        dt = C.c_double(0)
        self._ck(self.lib.titgpu_step(self.h, nsteps, C.byref(dt)), "titgpu_step")
        return dt.value

    def set_outputs(self, level):
        """0 state only, 1 + derived fields of fluid particles, 2 everything (reference, default)."""
        self._ck(self.lib.titgpu_set_outputs(self.h, int(level)), "titgpu_set_outputs")

    def set_lists(self, on):
        """Step-persistent candidate lists on / off (titgpu_set_lists)."""
        self._ck(self.lib.titgpu_set_lists(self.h, int(bool(on))), "titgpu_set_lists")

    def set_graphs(self, on):
        """Whole 2-D steps as CUDA graphs on / off (titgpu_set_graphs; default on)."""
        self._ck(self.lib.titgpu_set_graphs(self.h, int(bool(on))), "titgpu_set_graphs")

    @property
    def graph_replays(self):
        return int(self.lib.titgpu_graph_replays(self.h))

    def set_group_sweep(self, mode):
        """Grouped candidate sweep of the kernel-sum passes: -1 / None = by particle count (default), 0 = never, 1 = always (titgpu_set_group_sweep). Bit-identical results either way."""
        self._ck(self.lib.titgpu_set_group_sweep(self.h, -1 if mode is None else int(mode)), "titgpu_set_group_sweep")

    def set_tiles(self, on):
        """Shared-memory-staged kernel-sum pass on / off (titgpu_set_tiles; 3-D, radius-2h kernels)."""
        self._ck(self.lib.titgpu_set_tiles(self.h, int(bool(on))), "titgpu_set_tiles")

    @property
    def list_redos(self):
        return int(self.lib.titgpu_list_redos(self.h))

    # ---- slab decomposition (see slab.py) ----
    def mg_reserve(self, max_fluid):
        self._ck(self.lib.titgpu_mg_reserve(self.h, int(max_fluid)), "titgpu_mg_reserve")

    def mg_counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._ck(self.lib.titgpu_mg_counts(self.h, C.byref(a), C.byref(b), C.byref(c)), "titgpu_mg_counts")
        return a.value, b.value, c.value

    def mg_set_slab(self, axis, lo, hi, halo, fluid_total=-1):
        self._ck(self.lib.titgpu_mg_set_slab(self.h, int(axis), float(lo), float(hi), float(halo), int(fluid_total)), "titgpu_mg_set_slab")

    def mg_set_halo_pair(self, halo_pair):
        self._ck(self.lib.titgpu_mg_set_halo_pair(self.h, float(halo_pair)), "titgpu_mg_set_halo_pair")

    def mg_set_gids(self, gids):
        g = _arr(gids, np.int64)
        assert len(g) == self.n_fluid
        self._ck(self.lib.titgpu_mg_set_gids(self.h, g.ctypes.data), "titgpu_mg_set_gids")

    def mg_attach_nccl(self, id128: bytes, rank, nranks):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._ck(self.lib.titgpu_mg_attach_nccl(self.h, buf, int(rank), int(nranks)), "titgpu_mg_attach_nccl")

    def mg_attach_comm(self, comm_ptr, rank, nranks):
        self._ck(self.lib.titgpu_mg_attach_comm(self.h, comm_ptr, int(rank), int(nranks)), "titgpu_mg_attach_comm")

    def mg_attach_hub(self, hub, rank):
        self._ck(self.lib.titgpu_mg_attach_hub(self.h, hub, int(rank)), "titgpu_mg_attach_hub")

    def mg_detach(self):
        self._ck(self.lib.titgpu_mg_detach(self.h), "titgpu_mg_detach")

    def mg_download_owned(self, A_ptr=None, B_ptr=None, gid_ptr=None, cap=None):
        """Owned records into HOST buffers (raw addresses); returns n_owned. Without buffers: just the count."""
        n = C.c_size_t(0)
        self._ck(self.lib.titgpu_mg_download_owned(self.h, gid_ptr, A_ptr, B_ptr, int(cap if cap is not None else 0), C.byref(n)), "titgpu_mg_download_owned")
        return int(n.value)

    def mg_owned(self):
        """(gid, A, B) of the owned particles as numpy arrays (rank-local order)."""
        n = self.mg_counts()[0]
        gid, A, B = np.empty(max(n, 1), np.int64), np.empty((max(n, 1), 4)), np.empty((max(n, 1), 4))
        self.mg_download_owned(A.ctypes.data, B.ctypes.data, gid.ctypes.data, cap=max(n, 1))
        return gid[:n], A[:n], B[:n]

    def mg_upload_owned(self, n_owned, A_ptr, B_ptr, gid_ptr=None):
        self._ck(self.lib.titgpu_mg_upload_owned(self.h, int(n_owned), gid_ptr, A_ptr, B_ptr), "titgpu_mg_upload_owned")
        self.n_fluid = int(n_owned)

    def mg_stats(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self._ck(self.lib.titgpu_mg_stats(self.h, C.byref(a), C.byref(b)), "titgpu_mg_stats")
        return int(a.value), int(b.value)

    def _csr(self, fn, what):
        nnz = C.c_size_t(0)
        self._ck(fn(self.h, None, None, 0, C.byref(nnz)), what)
        off = np.zeros(self.n + 1, np.uint64)
        cols = np.zeros(max(nnz.value, 1), np.uint64)
        u64p = C.POINTER(C.c_uint64)
        self._ck(fn(self.h, off.ctypes.data_as(u64p), cols.ctypes.data_as(u64p), nnz.value, C.byref(nnz)), what)
        return off, cols[: nnz.value]

    def neighbors(self):
        """Sorted neighbour rows (CSR, self included): `mesh[a]`."""
        return self._csr(self.lib.titgpu_neighbors, "titgpu_neighbors")

    def face_neighbors(self):
        """Sorted rows of the domain faces meeting each particle's support sphere: `mesh[domain, a]`."""
        return self._csr(self.lib.titgpu_face_neighbors, "titgpu_face_neighbors")

    def synchronize(self):
        self._ck(self.lib.titgpu_synchronize(self.h), "titgpu_synchronize")

    def profile(self, on=True):
        self._ck(self.lib.titgpu_profile_enable(self.h, int(on)), "titgpu_profile_enable")

    def profile_reset(self):
        self._ck(self.lib.titgpu_profile_reset(self.h), "titgpu_profile_reset")

    def profile_read(self):
        """{kernel name: (launches, total device ms)} since the last reset."""
        out = {}
        for i in range(self.lib.titgpu_profile_count(self.h)):
            name, cnt, ms = C.c_char_p(), C.c_ulonglong(), C.c_double()
            self._ck(self.lib.titgpu_profile_get(self.h, i, C.byref(name), C.byref(cnt), C.byref(ms)), "titgpu_profile_get")
            out[name.value.decode()] = (int(cnt.value), float(ms.value))
        return out

    def measure_fp64_peak(self):
        t = C.c_double(0)
        self._ck(self.lib.titgpu_measure_fp64_peak(self.h, C.byref(t)), "titgpu_measure_fp64_peak")
        return t.value

    @property
    def launch_count(self):
        return int(self.lib.titgpu_launch_count(self.h))

    @property
    def stream(self):
        return self.lib.titgpu_stream(self.h)


def load_case(solver, case):
    """Feed a `cases.Case` into a solver, in the order of wcsph.cpp:79-154."""
    solver.set_params(case.g, case.mu, case.cs0, case.rho0, case.xi, case.h)
    solver.set_surface(case.verts, case.faces, case.cverts, case.cfaces)
    solver.set_particles(case.n_fluid, case.n_fixed)
    solver.upload("r", case.r)
    solver.upload("m", case.m)
    solver.upload("rho", case.rho)
