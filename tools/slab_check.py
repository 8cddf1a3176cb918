#!/usr/bin/env python3
"""Multi-rank parity check of the slab decomposition (run under torchrun, or
with --spawn N to start N ranks itself — several ranks may share one GPU over
gloo): every rank steps its slab, rank 0 gathers the owned particles and
compares them with a single-context run of the same case on its own GPU.

    torchrun --nproc-per-node 2 tools/slab_check.py --dim 3 --n-col 8 --steps 3
    python tools/slab_check.py --spawn 2 --backend gloo --dim 2 --n-col 24
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def run(rank, world, args):
    import titsolver_b200 as tb
    from titsolver_b200 import cases
    from titsolver_b200.slab import SlabSolver

    ndev = torch.cuda.device_count()
    device = rank % ndev
    torch.cuda.set_device(device)
    backend = args.backend or ("nccl" if ndev >= world else "gloo")
    if not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", device)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    case = cases.dam_break_2d(args.n_col) if args.dim == 2 else cases.dam_break_3d(args.n_col, wall_ratio=0.93, jitter=0.1)
    nf = case.n_fluid
    if args.kick:  # a velocity field that makes particles change slabs
        rng = np.random.default_rng(3)
        v0 = np.zeros_like(case.r)
        v0[:nf] = rng.normal(size=(nf, case.dim)) * args.kick
    ss = SlabSolver(case, rank, world, axis=args.axis, device=device)
    if args.kick:
        own = ss.gid.cpu().numpy()
        vl = np.zeros((len(own) + ss.n_fixed, case.dim))
        vl[: len(own)] = v0[own]
        ss.solver.upload("v", vl)
    ss.initialize()
    dts = [ss.step(1) for _ in range(args.steps)]
    gid, r, v, rho = ss.owned_state()
    parts = [None] * world
    dist.gather_object((gid.numpy(), r.numpy(), v.numpy(), rho.numpy(), ss.solver.mg_counts()), parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        g = np.concatenate([p[0] for p in parts])
        order = np.argsort(g)
        assert np.array_equal(g[order], np.arange(nf)), "ownership is not a partition of the fluid particles"
        R, V, RHO = (np.concatenate([p[k] for p in parts])[order] for k in (1, 2, 3))
        one = tb.Solver(case.dim, device=device)
        tb.load_case(one, case)
        if args.kick:
            one.upload("v", v0)
        one.initialize()
        dts1 = [one.step(1) for _ in range(args.steps)]
        res = {"world": world, "backend": backend, "dim": args.dim, "n": case.n, "steps": args.steps, "counts": [list(map(int, p[4])) for p in parts]}
        for name, a, b in (("r", R, one.download("r")[:nf]), ("v", V, one.download("v")[:nf]), ("rho", RHO, one.download("rho")[:nf])):
            err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
            res["err_" + name] = err
            ok = ok and err <= args.tol
        res["dt_err"] = float(max(abs(a - b) / b for a, b in zip(dts, dts1)))
        ok = ok and res["dt_err"] <= 1e-12
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def _spawned(rank, world, port, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.exit(run(rank, world, args))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--n-col", type=int, default=24)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--axis", type=int, default=0)
    ap.add_argument("--kick", type=float, default=0.0)
    ap.add_argument("--tol", type=float, default=1e-9)
    ap.add_argument("--backend", default="")
    ap.add_argument("--spawn", type=int, default=0)
    args = ap.parse_args()
    if args.spawn:
        import socket

        import torch.multiprocessing as mp

        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        ctx = mp.spawn(_spawned, args=(args.spawn, port, args), nprocs=args.spawn, join=True)
        return 0
    return run(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), args)


if __name__ == "__main__":
    sys.exit(main())
