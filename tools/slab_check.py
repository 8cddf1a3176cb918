#!/usr/bin/env python3
"""Multi-rank parity check of the slab decomposition: every rank steps its slab, the
owned particles are gathered and compared with a single-context run of the same case.

    torchrun --nproc-per-node 2 tools/slab_check.py --dim 3 --n-col 8 --steps 3   # one GPU per rank, NCCL
    python tools/slab_check.py --hub 3 --dim 2 --n-col 30                         # 3 ranks = 3 threads sharing the visible GPU(s)

Both modes run the same device-side exchange (csrc/mg.cuh); they differ in the transport
only (ncclSend/ncclRecv vs. event-ordered device copies inside one process).
"""
import argparse
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def make_case(args):
    from titsolver_b200 import cases

    if args.dim == 2:
        return cases.dam_break_2d(args.n_col)
    if args.lattice:
        return cases.dam_break_3d(args.n_col)
    return cases.dam_break_3d(args.n_col, wall_ratio=0.93, jitter=0.1)


def kick_velocity(case, kick):
    rng = np.random.default_rng(3)
    v0 = np.zeros_like(case.r)
    v0[: case.n_fluid] = rng.normal(size=(case.n_fluid, case.dim)) * kick
    return v0


def run_rank(case, rank, world, args, device, hub=None, v0=None):
    from titsolver_b200.slab import SlabSolver

    ss = SlabSolver(case, rank, world, axis=args.axis, device=device, hub=hub, integrator_id=args.integrator)
    if v0 is not None:
        ss.upload_owned_field("v", v0[ss.gid0])
    ss.initialize()
    dts = [ss.step(1) for _ in range(args.steps)]
    gid, r, v, rho = ss.owned_state()
    return (gid.numpy(), r.numpy(), v.numpy(), rho.numpy(), ss.solver.mg_counts(), ss.solver.mg_stats()), dts


def compare(case, parts, dts, args, device, backend, v0):
    import titsolver_b200 as tb

    nf = case.n_fluid
    g = np.concatenate([p[0] for p in parts])
    order = np.argsort(g)
    assert np.array_equal(g[order], np.arange(nf)), "ownership is not a partition of the fluid particles"
    R, V, RHO = (np.concatenate([p[k] for p in parts])[order] for k in (1, 2, 3))
    one = tb.Solver(case.dim, device=device, integrator_id=args.integrator)
    tb.load_case(one, case)
    if v0 is not None:
        one.upload("v", v0)
    one.initialize()
    dts1 = [one.step(1) for _ in range(args.steps)]
    res = {"world": len(parts), "backend": backend, "dim": args.dim, "n": case.n, "steps": args.steps, "counts": [list(map(int, p[4])) for p in parts],
           "exchanges_migrated": [list(map(int, p[5])) for p in parts]}
    ok = True
    for name, a, b in (("r", R, one.download("r")[:nf]), ("v", V, one.download("v")[:nf]), ("rho", RHO, one.download("rho")[:nf])):
        err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        res["err_" + name] = err
        ok = ok and err <= args.tol
    res["dt_err"] = float(max(abs(a - b) / b for a, b in zip(dts, dts1)))
    res["t_end"] = float(sum(dts1))
    ok = ok and res["dt_err"] <= 1e-12
    res["ok"] = bool(ok)
    print(json.dumps(res), flush=True)
    return ok


def run_hub(args):
    import titsolver_b200 as tb

    world = args.hub
    ndev = torch.cuda.device_count()
    case = make_case(args)
    v0 = kick_velocity(case, args.kick) if args.kick else None
    hub = tb.hub_create(world)
    out, errs = [None] * world, []

    def worker(rank):
        try:
            out[rank] = run_rank(case, rank, world, args, rank % ndev, hub=hub, v0=v0)
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        print(json.dumps({"ok": False, "errors": errs}), flush=True)
        return 1
    ok = compare(case, [o[0] for o in out], out[0][1], args, 0, "hub", v0)
    tb.hub_destroy(hub)
    return 0 if ok else 1


def run_dist(args):
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", str(rank)))
    if torch.cuda.device_count() < world:
        raise SystemExit("slab_check: NCCL needs one GPU per rank (use --hub N to share a GPU)")
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    case = make_case(args)
    v0 = kick_velocity(case, args.kick) if args.kick else None
    part, dts = run_rank(case, rank, world, args, device, v0=v0)
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        ok = compare(case, parts, dts, args, device, "nccl", v0)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--n-col", type=int, default=24)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--axis", type=int, default=0)
    ap.add_argument("--kick", type=float, default=0.0)
    ap.add_argument("--tol", type=float, default=1e-9)
    ap.add_argument("--integrator", type=int, default=3)
    ap.add_argument("--lattice", action="store_true", help="3-D: the benchmark lattice (wall spacing = particle spacing) instead of generic positions")
    ap.add_argument("--hub", type=int, default=0, help="run N ranks as threads of this process")
    args = ap.parse_args()
    return run_hub(args) if args.hub else run_dist(args)


if __name__ == "__main__":
    sys.exit(main())
