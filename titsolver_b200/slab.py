"""Slab domain decomposition of the WCSPH step across the GPUs of one box.

One process per GPU. Every rank owns the fluid particles of one slab
[lo, hi) along `axis` and keeps GHOST copies of the neighbouring slabs'
particles within a halo. The exchanges themselves (migration, halo set, the
refreshes before every neighbour search, {N, phi} and the shifted records
inside the shifting pass, the dt all-reduce) run INSIDE `titgpu_step` on the
context's stream: packing kernels + ncclSend / ncclRecv (csrc/mg.cuh,
csrc/mg_transport.cu). This module only cuts the case into slabs, hands every
rank's context its slab, halo width and communicator, and gathers results.

`exchange_records` / `NeighbourComm` below are the executable SPECIFICATION of
what one device-side exchange must deliver (owned + ghosts reproduce the global
neighbour sets; migration is a partition): pure torch, driven over gloo on CPU
by tests/test_slab_gloo.py.

This is the multi-GPU replacement of the reference's block partition
(/root/reference/source/tit/sph/particle_mesh.hpp:165-241), which splits the
particles among CPU threads with geom/partition/* and orders the pair loops so
that threads never touch the same particle; here slabs never touch the same
particle because every rank only updates what it owns.

Halo width (R = support radius, dw = largest wall-face edge): W = 2 R + dw plus a
margin for the motion within a step - an owned particle's wall terms read the
density of wall particles within R + dw, which is extrapolated from the fluid
within R of them (fluid_equations.hpp:122-164). Everything farther away that
the shifting / free-surface passes read of a neighbour (N, phi, the shifted
state: fluid_equations.hpp:409-414, 443-450, 489-511) is sent by its owner.

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


def rebalanced_edges(edges, costs, x_min, x_max, min_width, max_shift=None, ref_edges=None):
    """New slab edges that equalise a measured per-rank cost (e.g. the kernel time of the
    last steps). The cost of rank i is taken as uniformly spread over its slab - for the end
    slabs over the part that holds particles, [x_min, edges[1]) and [edges[-2], x_max] - and
    the new interior edges are the k / N quantiles of the resulting piecewise-linear
    cumulative cost. Every slab keeps at least `min_width` (the halo: ghosts come from
    adjacent slabs only) and no edge ends up farther than `max_shift` from `ref_edges` (default:
    `edges`) - what the rank-local wall pieces were cut for. The role of the reference's partitioners
    (geom/partition/sort_partition.hpp:25-75: equal COUNTS along a sort key) with measured
    cost in place of counts. Pure function: the same on every rank."""
    world = len(edges) - 1
    if world == 1:
        return list(edges)
    knots = [float(x_min)] + [float(e) for e in edges[1:-1]] + [float(x_max)]
    for a, b in zip(knots[:-1], knots[1:]):
        if not a < b:
            raise ValueError("rebalanced_edges: particles do not span every slab")
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(costs, dtype=np.float64))])
    if not cum[-1] > 0:
        return list(edges)
    targets = cum[-1] * np.arange(1, world) / world
    new = np.interp(targets, cum, knots)
    out = [-math.inf]
    for k in range(1, world):
        e = float(new[k - 1])
        if max_shift is not None:
            ref = (ref_edges or edges)[k]
            e = min(max(e, ref - max_shift), ref + max_shift)
        lo_prev = out[-1] if k > 1 else knots[0]
        e = max(e, lo_prev + min_width)
        out.append(e)
    # keep the tail feasible as well (walk back from the right end)
    hi_next = knots[-1]
    for k in range(world - 1, 0, -1):
        out[k] = min(out[k], hi_next - min_width)
        hi_next = out[k]
    out.append(math.inf)
    for a, b in zip(out[1:-2], out[2:-1]):
        if not b - a >= min_width * (1 - 1e-12):
            raise ValueError("rebalanced_edges: too many slabs for this extent")
    if not (out[1] > knots[0] and out[-2] < knots[-1]):
        raise ValueError("rebalanced_edges: too many slabs for this extent")
    return out


def halo_width(case, kernel_id=4, margin_dr=1.0):
    """(halo, R, dw): ghost-layer width 2 R + dw + margin, support radius, longest wall-face edge."""
    R = 2.0 * case.h if kernel_id not in (1, 2) else (2.5 if kernel_id == 1 else 3.0) * case.h
    dw = 0.0
    if len(case.faces):
        f = case.faces[: min(len(case.faces), 200000)].astype(np.int64)
        v = case.verts[f]
        for i in range(case.dim):
            d = v[:, i] - v[:, (i + 1) % case.dim]
            dw = max(dw, float(np.sqrt((d * d).sum(axis=1)).max()))
    return 2.0 * R + dw + margin_dr * case.dr, R, dw


class SlabSolver:
    """`titsolver_b200.Solver` of one rank, configured as one slab of a decomposed run.

    transport: `hub=<titsolver_b200.hub_create(world)>` for ranks that share a process
    (one thread each); otherwise NCCL - the ncclUniqueId made on rank 0 travels through
    torch.distributed (any backend), the communicator itself belongs to libtitgpu."""

    MARGIN_DR = 1.0  # extra halo, in particle spacings, for the motion within one step

    def __init__(self, case, rank, world, axis=0, edges=None, device=0, kernel_id=4, eos_id=0, integrator_id=3, group=None, reserve=None, local=False, hub=None,
                 fluid_total=None, thin_halo=True):
        """`case` is the GLOBAL case (every rank cuts out its slab; small runs and
        tests) or, with `local=True`, the rank-local one made by
        `cases.dam_break_3d_slab` (`edges` required, `case.meta["gid"]` = global ids)."""
        import titsolver_b200 as tb

        self.rank, self.world, self.axis = rank, world, axis
        self.dim = case.dim
        nf = case.n_fluid
        halo, R, dw = halo_width(case, kernel_id, self.MARGIN_DR)
        if not local:
            edges = edges or balanced_edges(case.r[:nf, axis], world)
            fluid_total = nf if fluid_total is None else fluid_total
        elif fluid_total is None:
            fluid_total = case.meta.get("n_fluid_global", -1)
        self.layout = SlabLayout(axis, list(edges), halo, halo)
        self.edges0 = list(edges)  # the edges the rank-local wall pieces were cut for
        lo, hi = self.layout.bounds(rank)
        self.lo, self.hi = lo, hi
        if not local:
            x = case.r[:nf, axis]
            own = np.nonzero((x >= lo) & (x < hi))[0]
            wf = halo + R + 2 * dw
            lverts, lfaces, vid = local_surface(case.verts, case.faces, axis, lo - wf, hi + wf)
            r = np.concatenate([case.r[own], lverts], axis=0)
            mass = np.concatenate([case.m[own], case.m[nf + vid]])
            rho = np.concatenate([case.rho[own], case.rho[nf + vid]])
            gid = own.astype(np.int64)
            if reserve is None:  # owned + the particles now inside the halo, with head-room
                reserve = int(1.3 * np.count_nonzero((x >= lo - halo) & (x < hi + halo))) + 4096
        else:
            lverts, lfaces, r, mass, rho, gid = case.verts, case.faces, case.r, case.m, case.rho, np.asarray(case.meta["gid"], dtype=np.int64)
            if reserve is None:  # both halos at the slab's particle density
                x = case.r[:nf, axis]
                width = max(float(x.max() - x.min()) + case.dr, case.dr) if nf else case.dr
                reserve = int(1.25 * nf * (1.0 + 2.0 * halo / width)) + 4096
        self.n_fixed = len(lverts)
        self.gid0 = gid  # global ids of the initially owned particles, in upload order
        self.solver = s = tb.Solver(case.dim, kernel_id, eos_id, integrator_id, device=device)
        s.set_params(case.g, case.mu, case.cs0, case.rho0, case.xi, case.h)
        s.set_surface(lverts, lfaces, case.cverts, case.cfaces)
        n_own = len(gid)
        s.mg_reserve(max(int(reserve), n_own))
        s.set_particles(n_own, self.n_fixed)
        s.upload("r", r)
        s.upload("m", mass)
        s.upload("rho", rho)
        s.mg_set_gids(gid)
        self.fluid_total = fluid_total
        s.mg_set_slab(axis, lo, hi, halo, fluid_total)
        if thin_halo:  # one support radius + margin away from the walls, the full halo along them
            s.mg_set_halo_pair(R + self.MARGIN_DR * case.dr)
        if world > 1:
            if hub is not None:
                s.mg_attach_hub(hub, rank)
            else:
                box = [tb.nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(box, src=0, group=group)
                s.mg_attach_nccl(box[0], rank, world)

    def rebalance(self, cost, extent, max_shift=None, group=None):
        """Equalise `cost` (this rank's measured time per step, any unit) over the ranks by moving
        the slab edges (`rebalanced_edges`); the next step migrates the particles that changed
        slab. `extent` = (x_min, x_max) of the fluid along the slab axis, the same on every rank.
        Collective over torch.distributed (all ranks call it with their own cost)."""
        costs = [None] * self.world
        dist.all_gather_object(costs, float(cost), group=group)
        edges = rebalanced_edges(self.layout.edges, costs, extent[0], extent[1], self.layout.w_rhs, max_shift, self.edges0)
        self.set_edges(edges)
        return edges

    def set_edges(self, edges):
        self.layout.edges = list(edges)
        self.lo, self.hi = self.layout.bounds(self.rank)
        self.solver.mg_set_slab(self.axis, self.lo, self.hi, self.layout.w_rhs, self.fluid_total)

    def upload_owned_field(self, field, values):
        """Set `field` ("v", "rho", ...) of the owned particles before the first step
        (`values` in this rank's owned order; wall particles keep zeros / their values)."""
        s = self.solver
        full = s.download(field)
        full[: s.n_fluid] = values
        s.upload(field, full)

    # ---- user surface ----
    def initialize(self):
        self.solver.initialize()

    def step(self, nsteps=1):
        return self.solver.step(nsteps)

    def owned_state(self):
        """(gid, r, v, rho) of the particles this rank owns, as CPU tensors."""
        gid, A, B = self.solver.mg_owned()
        gid, A, B = torch.from_numpy(gid.copy()), torch.from_numpy(A.copy()), torch.from_numpy(B.copy())
        if self.dim == 3:
            return gid, A[:, 0:3], B[:, 0:3], A[:, 3]
        return gid, A[:, 0:2], B[:, 0:2], A[:, 2]

    @property
    def n_owned(self):
        return self.solver.mg_counts()[0]
