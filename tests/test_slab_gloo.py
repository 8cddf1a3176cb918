"""World-size-2 and -3 gloo tests (CPU) of the slab decomposition's host logic:
`titsolver_b200.slab.exchange_records` / `balanced_edges` / `local_surface` —
the very code the GPU path runs over NCCL, here on CPU tensors. The oracle
provides the neighbour sets: after a halo exchange every owned particle must
find, among owned + ghost particles, exactly its global neighbours."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cloud(n=700, dim=2, seed=5):
    rng = np.random.default_rng(seed)
    r = rng.uniform(0.0, 1.0, size=(n, dim))
    r[:, 0] *= 3.0  # long axis = slab axis
    return r


def _records(r, v=None):
    """Pack like the 2-D / 3-D device records: A = {x, y, (z,) rho..}, B = {v.., m}."""
    n, dim = r.shape
    A = np.zeros((n, 4)); B = np.zeros((n, 4))
    A[:, :dim] = r
    A[:, 3 if dim == 3 else 2] = 1000.0 + np.arange(n)
    if v is not None:
        B[:, :dim] = v
    return torch.from_numpy(np.concatenate([A, B], axis=1))


def _neighbour_sets(oracle_lib, r, radius, ids):
    s = oracle_lib.OracleSolver(r.shape[1], 0)
    s.set_params(9.81, 1e-3, 10.0, 1000.0, 7.0, radius / 2.0)
    e = np.zeros((0, r.shape[1]))
    s.set_surface(e, e.astype(np.uint64), e, e.astype(np.uint64))
    s.set_particles(len(r), 0)
    s.upload("r", r)
    off, cols = s.neighbors()
    return [set(ids[cols[int(off[i]):int(off[i + 1])].astype(np.int64)].tolist()) for i in range(len(r))]


def _worker(rank, world, port, dim, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import oracle_lib
    from titsolver_b200.slab import NeighbourComm, balanced_edges, exchange_records

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        radius = 0.11
        r = _cloud(dim=dim)
        n = len(r)
        edges = balanced_edges(r[:, 0], world)
        lo, hi = edges[rank], edges[rank + 1]
        comm = NeighbourComm(rank, world)
        own = np.nonzero((r[:, 0] >= lo) & (r[:, 0] < hi))[0]
        assert abs(len(own) - n / world) <= 0.05 * n + 2  # balanced
        rec, gid = _records(r[own]), torch.from_numpy(own.astype(np.int64))

        # (1) static halo exchange: owned + ghosts reproduce the global neighbour sets
        rec1, gid1, n_owned = exchange_records(rec, gid, len(own), lo, hi, radius, True, 0, comm)
        assert n_owned == len(own) and torch.equal(gid1[:n_owned], gid)
        g = gid1.numpy()
        assert len(set(g.tolist())) == len(g)  # no duplicates
        ghosts = r[g[n_owned:], 0]
        assert np.all((ghosts < lo) | (ghosts >= hi))
        assert np.all((ghosts >= lo - radius) & (ghosts < hi + radius))
        local = _neighbour_sets(oracle_lib, rec1[:, :dim].numpy().copy(), radius, g)
        glob = _neighbour_sets(oracle_lib, r, radius, np.arange(n))
        for i in range(n_owned):
            assert local[i] == glob[g[i]], (rank, i)

        # (2) motion + migration: each particle moves by a gid-seeded displacement < radius / 2
        rng = np.random.default_rng(11)
        disp = rng.uniform(-0.5, 0.5, size=(n, dim)) * radius * 0.9
        r2 = r + disp
        rec_m = rec1.clone()
        rec_m[:, :dim] = torch.from_numpy(r2[g])        # ghosts are stale copies: they are dropped anyway
        rec2, gid2, n_owned2 = exchange_records(rec_m, gid1, n_owned, lo, hi, 2 * radius, True, 0, comm)
        g2 = gid2.numpy()
        x_owned = rec2[:n_owned2, 0].numpy()
        assert np.all((x_owned >= lo) & (x_owned < hi))
        assert np.allclose(rec2[:, :dim].numpy(), r2[g2])  # records travelled with their ids
        owners = [None] * world
        dist.all_gather_object(owners, g2[:n_owned2].tolist())
        allg = sorted(sum(owners, []))
        assert allg == list(range(n)), "every particle has exactly one owner after migration"
        local = _neighbour_sets(oracle_lib, rec2[:, :dim].numpy().copy(), radius, g2)
        glob = _neighbour_sets(oracle_lib, r2, radius, np.arange(n))
        for i in range(n_owned2):
            assert local[i] == glob[g2[i]], (rank, i)

        # (3) halo refresh inside a step: ownership frozen, u_old columns travel with the owned rows only
        rec16 = torch.cat([rec2, rec2 + 0.5], dim=1)
        rec3, gid3, n_owned3 = exchange_records(rec16, gid2, n_owned2, lo, hi, radius, False, 0, comm)
        assert n_owned3 == n_owned2 and torch.equal(rec3[:n_owned3], rec16[:n_owned2]) and torch.equal(gid3[:n_owned3], gid2[:n_owned2])
        assert rec3.shape[1] == 16 and float(rec3[n_owned3:, 8:].abs().sum()) == 0.0

        # (4) the time-step reduction
        tmin, tmax = torch.tensor([1.0 + rank]), torch.tensor([10.0 + rank])
        comm.allreduce_min_max(tmin, tmax)
        assert tmin.item() == 1.0 and tmax.item() == 10.0 + world - 1
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim", [(2, 2), (3, 3)])
def test_slab_exchange_over_gloo(world, dim, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, dim, q)) for rk in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rk, msg in res:
        assert msg == "ok", f"rank {rk}: {msg}"


def test_local_surface_subset():
    from titsolver_b200 import cases
    from titsolver_b200.slab import local_surface

    case = cases.dam_break_2d(10)
    v, f, vid = local_surface(case.verts, case.faces, 0, 0.5, 1.0)
    assert len(f) > 0 and np.array_equal(v, case.verts[vid])
    x = v[f.astype(np.int64), 0]
    assert np.all((x.max(axis=1) >= 0.5) & (x.min(axis=1) <= 1.0))
    # the subset keeps the geometry of every selected face
    sel = (case.verts[case.faces.astype(np.int64), 0].max(axis=1) >= 0.5) & (case.verts[case.faces.astype(np.int64), 0].min(axis=1) <= 1.0)
    assert np.array_equal(v[f.astype(np.int64)], case.verts[case.faces[sel].astype(np.int64)])
