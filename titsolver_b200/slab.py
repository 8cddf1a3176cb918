"""Slab domain decomposition of the WCSPH step across the GPUs of one box.

One process per GPU. Every rank owns the fluid particles of one slab
[lo, hi) along `axis` and, before every neighbour search, receives GHOST copies
of the neighbouring slabs' particles within a halo (NCCL send/recv over NVLink
through torch.distributed; gloo on CPU for the tests). Ghosts are neighbours
only. Ownership changes once per step, at its first search. The time-step
scalars are reduced over the ranks once per step (MIN / MAX all-reduce).

This is the multi-GPU replacement of the reference's block partition
(/root/reference/source/tit/sph/particle_mesh.hpp:165-241), which splits the
particles among CPU threads with geom/partition/* and orders the pair loops so
that threads never touch the same particle; here slabs never touch the same
particle because every rank only updates what it owns.

Halo widths (R = support radius, dw = largest wall-face edge):
  right-hand-side searches  W = 2 R + dw   (a neighbour's wall density needs the
                                            fluid within R of a wall vertex that
                                            is within R + dw of an owned particle)
  post_integrate search     W = 4 R + dw   (free-surface correction reads shifted
                                            neighbours, whose shift reads N / phi
                                            of their neighbours, whose sums read
                                            wall densities: fluid_equations.hpp:331-512)
plus a margin for the motion within a step.

`exchange_records` and `SlabLayout` are pure torch and device-agnostic: the
CPU tests drive exactly this code over gloo.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class SlabLayout:
    axis: int
    edges: list  # world + 1 increasing positions; the outer two are -inf / +inf
    w_rhs: float
    w_post: float

    def bounds(self, rank):
        return self.edges[rank], self.edges[rank + 1]

    @property
    def world(self):
        return len(self.edges) - 1


def balanced_edges(x_fluid: np.ndarray, world: int) -> list:
    """Slab edges with about equal numbers of fluid particles. Edges are placed
    half-way between distinct particle planes so that no particle of the initial
    lattice sits on an edge."""
    xs = np.unique(x_fluid)
    counts = np.searchsorted(np.sort(x_fluid), xs, side="right")  # cumulative count up to each plane
    n = len(x_fluid)
    edges = [-math.inf]
    for k in range(1, world):
        i = int(np.searchsorted(counts, k * n / world, side="left"))
        i = min(max(i, 0), len(xs) - 2)
        edges.append(0.5 * (float(xs[i]) + float(xs[i + 1])))
    edges.append(math.inf)
    for a, b in zip(edges[:-1], edges[1:]):
        if not a < b:
            raise ValueError("too many slabs for this particle set")
    return edges


class NeighbourComm:
    """Point-to-point exchange with the left / right slab (torch.distributed)."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        # gloo moves CPU tensors only: device payloads are staged through the host
        # (tests with several ranks on one GPU); NCCL sends device memory directly.
        self.host_staged = world > 1 and dist.is_initialized() and dist.get_backend(group) == "gloo"

    def sendrecv(self, to_left: torch.Tensor, to_right: torch.Tensor):
        """Send (k, c) float64 payloads to both neighbours, return what they sent."""
        out_dev = to_left.device
        if self.host_staged and to_left.is_cuda:
            to_left, to_right = to_left.cpu(), to_right.cpu()
        dev, dt, c = to_left.device, to_left.dtype, to_left.shape[1]
        peers = [(self.left, to_left), (self.right, to_right)]
        # 1) row counts
        ops, cnt_in = [], {}
        for peer, payload in peers:
            if peer is None:
                continue
            cnt_in[peer] = torch.zeros(1, dtype=torch.int64, device=dev)
            ops.append(dist.P2POp(dist.isend, torch.tensor([payload.shape[0]], dtype=torch.int64, device=dev), peer, self.group))
            ops.append(dist.P2POp(dist.irecv, cnt_in[peer], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # 2) payloads
        ops, bufs = [], {}
        for peer, payload in peers:
            if peer is None:
                continue
            bufs[peer] = torch.empty((int(cnt_in[peer].item()), c), dtype=dt, device=dev)
            if payload.shape[0]:
                ops.append(dist.P2POp(dist.isend, payload.contiguous(), peer, self.group))
            if bufs[peer].shape[0]:
                ops.append(dist.P2POp(dist.irecv, bufs[peer], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        empty = torch.empty((0, c), dtype=dt, device=dev)
        return bufs.get(self.left, empty).to(out_dev), bufs.get(self.right, empty).to(out_dev)

    def allreduce_min_max(self, t_min: torch.Tensor, t_max: torch.Tensor):
        if self.world > 1:
            if self.host_staged and t_min.is_cuda:
                a, b = t_min.cpu(), t_max.cpu()
                dist.all_reduce(a, op=dist.ReduceOp.MIN, group=self.group)
                dist.all_reduce(b, op=dist.ReduceOp.MAX, group=self.group)
                t_min.copy_(a); t_max.copy_(b)
                return
            dist.all_reduce(t_min, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX, group=self.group)


def exchange_records(rec: torch.Tensor, gid: torch.Tensor, n_owned: int, lo: float, hi: float, halo: float, migrate: bool, axis: int, comm: NeighbourComm):
    """One halo exchange (+ ownership migration when `migrate`).

    rec   (nf, 8 | 16) float64 rank-local fluid records [A | B | A0 | B0], owned rows first
    gid   (nf,) int64 global particle ids
    Returns (rec', gid', n_owned'): owned rows first, then the new ghosts (their
    A0 | B0 columns, if any, are zero: ghosts are never updated).
    """
    cols = rec.shape[1]
    own, gown = rec[:n_owned], gid[:n_owned]
    x = own[:, axis]
    has_l, has_r = comm.left is not None, comm.right is not None
    none = torch.zeros_like(x, dtype=torch.bool)
    send_l = (x < lo + halo) if has_l else none
    send_r = (x >= hi - halo) if has_r else none
    out = (((x < lo) if has_l else none) | ((x >= hi) if has_r else none)) if migrate else none

    def payload(mask):
        idx = torch.nonzero(mask).squeeze(1)
        return torch.cat([own[idx, :8], gown[idx].to(rec.dtype).unsqueeze(1)], dim=1)

    from_l, from_r = comm.sendrecv(payload(send_l), payload(send_r))
    got = torch.cat([from_l, from_r], dim=0)
    got_rec, got_gid = got[:, :8], got[:, 8].to(torch.int64)
    if cols > 8:
        got_rec = torch.cat([got_rec, torch.zeros((got_rec.shape[0], cols - 8), dtype=rec.dtype, device=rec.device)], dim=1)
    if migrate:
        gx = got_rec[:, axis]
        mine = (gx >= lo) & (gx < hi)
        keep = ~out
        new_own = torch.cat([own[keep], got_rec[mine]], dim=0)
        new_gown = torch.cat([gown[keep], got_gid[mine]], dim=0)
        ghosts = torch.cat([own[out], got_rec[~mine]], dim=0)
        ggid = torch.cat([gown[out], got_gid[~mine]], dim=0)
    else:
        new_own, new_gown, ghosts, ggid = own, gown, got_rec, got_gid
    return torch.cat([new_own, ghosts], dim=0), torch.cat([new_gown, ggid], dim=0), int(new_own.shape[0])


def local_surface(verts: np.ndarray, faces: np.ndarray, axis: int, lo: float, hi: float):
    """Faces whose extent along `axis` meets [lo, hi], with their vertices renumbered.
    Returns (local verts, local faces, global ids of the local vertices)."""
    if len(faces) == 0:
        return verts[:0], faces[:0], np.zeros(0, np.int64)
    fx = verts[faces.astype(np.int64), axis]  # (nf, dim)
    sel = (fx.max(axis=1) >= lo) & (fx.min(axis=1) <= hi)
    f = faces[sel].astype(np.int64)
    vid, inv = np.unique(f.ravel(), return_inverse=True)
    return verts[vid], inv.reshape(f.shape).astype(np.uint64), vid


class _DevicePtr:
    """A raw CUDA pointer as a torch-importable array (`__cuda_array_interface__`)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class SlabSolver:
    """`titsolver_b200.Solver` of one rank + the halo exchange around it."""

    MARGIN_DR = 1.0  # extra halo, in particle spacings, for the motion within one step

    def __init__(self, case, rank, world, axis=0, edges=None, device=0, kernel_id=4, eos_id=0, integrator_id=3, group=None, reserve=None, local=False):
        """`case` is the GLOBAL case (every rank cuts out its slab; small runs and
        tests) or, with `local=True`, the rank-local one made by
        `cases.dam_break_3d_slab` (`edges` required, `case.meta["gid"]` = global ids)."""
        import titsolver_b200 as tb

        self.rank, self.world, self.axis = rank, world, axis
        self.comm = NeighbourComm(rank, world, group)
        self.dim = case.dim
        nf = case.n_fluid
        R = 2.0 * case.h if kernel_id not in (1, 2) else (2.5 if kernel_id == 1 else 3.0) * case.h
        dw = self._max_face_edge(case)
        m = self.MARGIN_DR * case.dr
        if not local:
            edges = edges or balanced_edges(case.r[:nf, axis], world)
        self.layout = SlabLayout(axis, edges, 2 * R + dw + m, 4 * R + dw + m)
        lo, hi = self.layout.bounds(rank)
        self.lo, self.hi = lo, hi
        if not local:
            x = case.r[:nf, axis]
            own = np.nonzero((x >= lo) & (x < hi))[0]
            wf = self.layout.w_post + R + 2 * dw
            lverts, lfaces, vid = local_surface(case.verts, case.faces, axis, lo - wf, hi + wf)
            r = np.concatenate([case.r[own], lverts], axis=0)
            mass = np.concatenate([case.m[own], case.m[nf + vid]])
            rho = np.concatenate([case.rho[own], case.rho[nf + vid]])
            gid = own.astype(np.int64)
            if reserve is None:  # owned + the particles now inside the widest halo, with head-room
                w = self.layout.w_post
                reserve = int(1.3 * np.count_nonzero((x >= lo - w) & (x < hi + w))) + 4096
        else:
            lverts, lfaces, r, mass, rho, gid = case.verts, case.faces, case.r, case.m, case.rho, np.asarray(case.meta["gid"], dtype=np.int64)
            if reserve is None:  # both halos at the slab's particle density
                x = case.r[:nf, axis]
                width = max(float(x.max() - x.min()) + case.dr, case.dr)
                reserve = int(1.3 * nf * (1.0 + 2.0 * self.layout.w_post / width)) + 4096
        self.n_fixed = len(lverts)
        self.solver = s = tb.Solver(case.dim, kernel_id, eos_id, integrator_id, device=device)
        s.set_params(case.g, case.mu, case.cs0, case.rho0, case.xi, case.h)
        s.set_surface(lverts, lfaces, case.cverts, case.cfaces)
        n_own = len(gid)
        s.mg_reserve(max(int(reserve), n_own))
        s.set_particles(n_own, self.n_fixed)
        s.upload("r", r)
        s.upload("m", mass)
        s.upload("rho", rho)
        self.torch_device = torch.device("cuda", device)
        self.gid = torch.from_numpy(gid).to(self.torch_device)
        self._stream = torch.cuda.ExternalStream(s.stream, device=self.torch_device)
        self._exc = None
        self._bufs, self._buf_cap = [], 0
        self.fast_ghost_refresh = True
        self.exchange_ms = 0.0
        s.mg_set_exchange(self._on_exchange)

    @staticmethod
    def _max_face_edge(case):
        if len(case.faces) == 0:
            return 0.0
        f = case.faces[: min(len(case.faces), 200000)].astype(np.int64)
        v = case.verts[f]
        e = 0.0
        for i in range(case.dim):
            d = v[:, i] - v[:, (i + 1) % case.dim]
            e = max(e, float(np.sqrt((d * d).sum(axis=1)).max()))
        return e

    # ---- the exchange callback (invoked from inside titgpu_step) ----
    def _on_exchange(self, phase):
        try:
            with torch.cuda.stream(self._stream):
                if phase == 2:
                    sc = torch.as_tensor(_DevicePtr(self.solver.mg_scalars, 3), device=self.torch_device)
                    self.comm.allreduce_min_max(sc[2:3], sc[1:2])
                else:
                    self._halo(migrate=(phase == 0), halo=self.layout.w_post if phase == 3 else self.layout.w_rhs, with_old=(phase != 0 and self.solver_integrator_has_old))
            return 0
        except Exception as e:  # surfaces through titgpu_step's status
            self._exc = e
            return 1

    @property
    def solver_integrator_has_old(self):
        return True

    def _export(self, with_old):
        n_owned, n_ghost, _ = self.solver.mg_counts()
        nf = n_owned + n_ghost
        bufs = [torch.empty((max(nf, 1), 4), dtype=torch.float64, device=self.torch_device) for _ in range(4 if with_old else 2)]
        ptrs = [b.data_ptr() for b in bufs] + [None] * (4 - len(bufs))
        self.solver.mg_export(*ptrs)
        return torch.cat([b[:nf] for b in bufs], dim=1), n_owned

    def _halo_ghosts(self, halo, with_old):
        """Refresh of the ghost layer without a change of ownership (every exchange of a
        step but the first): the owned records never leave the export buffers. They are
        exported once, the rows within `halo` of a slab edge are sent, the received
        ghosts are written behind the owned rows of the same buffers and those are
        imported back - two passes over the rank's records instead of the six of the
        general path (concatenations, boolean gathers)."""
        s = self.solver
        n_owned, n_ghost, _ = s.mg_counts()
        nf = n_owned + n_ghost
        k = 4 if with_old else 2
        cap = max(self._buf_cap, nf)
        if len(self._bufs) < k or self._bufs[0].shape[0] < cap:
            cap = int(cap * 1.1) + 1024
            self._bufs = [torch.empty((cap, 4), dtype=torch.float64, device=self.torch_device) for _ in range(4)]
            self._buf_cap = cap
        bufs = self._bufs[:k]
        s.mg_export(*([b.data_ptr() for b in bufs] + [None] * (4 - k)))
        A, B = bufs[0], bufs[1]
        x = A[:n_owned, self.axis]
        gown = self.gid[:n_owned]
        comm = self.comm

        def payload(mask):
            idx = torch.nonzero(mask).squeeze(1)
            return torch.cat([A[idx], B[idx], gown[idx].to(torch.float64).unsqueeze(1)], dim=1)

        empty = torch.empty((0, 9), dtype=torch.float64, device=self.torch_device)
        to_l = payload(x < self.lo + halo) if comm.left is not None else empty
        to_r = payload(x >= self.hi - halo) if comm.right is not None else empty
        from_l, from_r = comm.sendrecv(to_l, to_r)
        got = torch.cat([from_l, from_r], dim=0)
        ng = int(got.shape[0])
        if n_owned + ng > self._bufs[0].shape[0]:
            grown = [torch.empty((int((n_owned + ng) * 1.1) + 1024, 4), dtype=torch.float64, device=self.torch_device) for _ in range(4)]
            for g, b in zip(grown, self._bufs):
                g[:n_owned] = b[:n_owned]
            self._bufs, self._buf_cap = grown, grown[0].shape[0]
            bufs = self._bufs[:k]
            A, B = bufs[0], bufs[1]
        A[n_owned:n_owned + ng] = got[:, 0:4]
        B[n_owned:n_owned + ng] = got[:, 4:8]
        for b in bufs[2:]:
            b[n_owned:n_owned + ng] = 0.0  # ghosts are never updated: no old state
        s.mg_import(n_owned, ng, *([b.data_ptr() if n_owned + ng else None for b in bufs] + [None] * (4 - k)))
        self.gid = torch.cat([gown, got[:, 8].to(torch.int64)])

    def _halo(self, migrate, halo, with_old):
        if not migrate and self.fast_ghost_refresh:
            return self._halo_ghosts(halo, with_old)
        rec, n_owned = self._export(with_old)
        rec2, gid2, n_owned2 = exchange_records(rec, self.gid, n_owned, self.lo, self.hi, halo, migrate, self.axis, self.comm)
        nf2 = rec2.shape[0]
        parts = [rec2[:, 4 * k: 4 * k + 4].contiguous() for k in range(rec2.shape[1] // 4)]
        ptrs = [p.data_ptr() if nf2 else None for p in parts] + [None] * (4 - len(parts))
        self.solver.mg_import(n_owned2, nf2 - n_owned2, *ptrs)
        self.gid = gid2
        self._keep = parts  # alive until the next exchange (the import is asynchronous)

    # ---- user surface ----
    def initialize(self):
        self.solver.initialize()

    def step(self, nsteps=1):
        try:
            return self.solver.step(nsteps)
        except Exception:
            if self._exc is not None:
                exc, self._exc = self._exc, None
                raise exc
            raise

    def owned_state(self):
        """(gid, r, v, rho) of the particles this rank owns, as CPU tensors."""
        with torch.cuda.stream(self._stream):
            rec, n_owned = self._export(False)
            rec, gid = rec[:n_owned].cpu(), self.gid[:n_owned].cpu()
        D = self.dim
        if D == 3:
            return gid, rec[:, 0:3], rec[:, 4:7], rec[:, 3]
        return gid, rec[:, 0:2], rec[:, 4:6], rec[:, 2]

    @property
    def n_owned(self):
        return self.solver.mg_counts()[0]
