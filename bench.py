#!/usr/bin/env python
"""bench.py — WCSPH particle-updates/s of the B200 path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c1|c2|c3|c4|c5] [--n-col N]

One "step" = one call of SSPRKIntegrator::step (order three) over the whole
particle set (/root/reference/source/tit/sph/time_integrator.hpp:161-184) =
`n` particle-updates. The default workload is BASELINE.json configs[2], the
single-GPU configuration its metric is quoted on: the 3-D dam break with
~10 M fluid particles (+ ~1.8 M wall particles), fp64, Wendland C4, Tait EOS.

Prints ONE JSON line (rank 0). `value` times K steps with the state resident in
HBM; `e2e` times K steps through the C ABI with HOST buffers (upload r, v, rho
from pinned memory, step, download r, v, rho every step). `roofline` is the
kernel-sum pass (k_rhs) against the measured HBM peak, with the FP64-pipe
figures beside it; `cpu_baseline` is the CPU restatement of the reference
(oracle/, OpenMP, all host cores) on a bounded sample.

`--impl reference` times the CPU restatement alone (the reference itself needs
C++26 / GCC 16 + oneTBB and cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "WCSPH particle-updates/s"
UNIT = "particle-updates/s"

# fluid lattice columns per H
WORKLOADS = {
    "c1": dict(dim=2, n_col=80, label="2D dam break 160x80 (titwcsph default case)"),
    "c2": dict(dim=2, n_col=707, label="2D dam break 1414x707 (~1M particles)"),
    "c3": dict(dim=3, n_col=171, label="3D dam break 342x171x170 (~10M fluid + 1.8M wall particles)"),
    "c4": dict(dim=3, n_col=272, label="3D dam break 544x272x271 (~40M fluid particles)"),
    "c5": dict(dim=3, n_col=368, label="3D dam break 736x368x367 (~100M fluid particles)"),
}


def make_case(dim, n_col):
    from titsolver_b200 import cases

    return cases.dam_break_2d(n_col) if dim == 2 else cases.dam_break_3d(n_col)


def alg_bytes_rhs(dim):
    """Algorithmic bytes of one kernel-sum (fused RHS + update) pass per particle:
    read r, v, rho, m, gamma; write r', v', rho' = 4V + 4S."""
    return 4 * 8 * dim + 4 * 8


def alg_bytes_step(dim):
    """B_alg = 52V + 2T + 40S + 64 bytes per particle-update."""
    V, T, S = 8 * dim, 8 * dim * dim, 8
    return 52 * V + 2 * T + 40 * S + 64


def alg_flops_rhs(dim):
    """Gather-form flops of one kernel-sum pass per particle: K neighbours x fused RHS."""
    return (48 * 72) if dim == 2 else (256 * 90)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(dim, n):
    """DRAM bytes (read + write) of one k_rhs launch over n particles, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json: bytes per particle measured at the
    capture's size, scaled to this launch); None when no capture exists for this dimension."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f).get("k_rhs", {}).get(f"{dim}d")
    except (OSError, ValueError):
        return None, None
    if not rec:
        return None, None
    return rec["bytes_per_particle"] * n, rec["source"]


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_sample(dim, budget_s=15.0, n_col=None):
    """Time the CPU restatement (oracle, -O3 OpenMP build) on a bounded sample of
    the same workload; returns the cpu_baseline object."""
    import oracle_lib

    if n_col is None:
        n_col = 200 if dim == 2 else 28
    case = make_case(dim, n_col)
    s = oracle_lib.OracleSolver(dim, fast=True)
    s.lib.orc_set_num_threads(host_threads())
    oracle_lib.load_case(s, case)
    s.initialize()
    s.step(1)  # warm-up (first-touch, thread pool)
    steps, t0 = 0, time.perf_counter()
    while True:
        s.step(1)
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or steps >= 50:
            break
    cores = int(s.lib.orc_num_threads())
    return {
        "value": case.n * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{steps} SSPRK3 steps of the {dim}D dam break at n_col={n_col} ({case.n} particles, {100.0 * case.n_fixed / case.n:.0f} % of them "
                  f"wall particles: the closed tank at a size the CPU finishes, so the wall integrals weigh more than at the benchmark size) "
                  f"in {el:.1f} s, oracle/liboracle_fast.so (-O3, OpenMP, gather form)",
    }, case.n, steps, el


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    dim = w["dim"]
    # Each "step" is a bounded sample: one SSPRK3 step of the reduced-size case.
    import oracle_lib

    n_col = args.ref_n_col or (200 if dim == 2 else 28)
    case = make_case(dim, n_col)
    s = oracle_lib.OracleSolver(dim, fast=True)
    s.lib.orc_set_num_threads(host_threads())  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    oracle_lib.load_case(s, case)
    s.initialize()
    for _ in range(args.warmup):
        s.step(1)
    t0 = time.perf_counter()
    s.step(args.steps)
    el = time.perf_counter() - t0
    val = case.n * args.steps / el
    cores = int(s.lib.orc_num_threads())
    sample = f"{args.steps} SSPRK3 steps of the {dim}D dam break at n_col={n_col} ({case.n} particles) per run; throughput per particle-update"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "sample_n_col": n_col, "sample_particles": case.n, "integrator": "ssprk3", "kernel": "SixthOrderWendland", "eos": "tait"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference (OpenMP); the reference itself (C++26, oneTBB) cannot be built in this image",
    }
    emit(json.dumps(out))


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch

    import titsolver_b200 as tb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    w = WORKLOADS[args.workload]
    dim = w["dim"]
    n_col = args.n_col or w["n_col"]
    slab = None
    if world > 1:
        # Weak scaling: the 3-D tank is made `world` times deeper along z and cut into
        # `world` slabs (one per GPU) with ghost layers exchanged over NCCL send/recv
        # before every neighbour search (titsolver_b200/slab.py).
        if dim != 3:
            raise SystemExit("bench.py: the multi-GPU workload is the 3-D dam break (use --workload c3/c4/c5)")
        from titsolver_b200 import cases
        from titsolver_b200.slab import SlabSolver

        case, edges = cases.dam_break_3d_slab(n_col, world, rank)
        slab = SlabSolver(case, rank, world, axis=2, edges=edges, device=local_rank, local=True)
        solver = slab.solver
        n_global = case.meta["n_fluid_global"] + case.meta["n_fixed_global"]
        n = n_global / world  # particles per GPU (wall particles of the halos are not counted twice)
    else:
        case = make_case(dim, n_col)
        n = case.n
        solver = tb.Solver(dim, device=local_rank)
        tb.load_case(solver, case)
    solver.initialize()
    stream = torch.cuda.ExternalStream(solver.stream, device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        solver.synchronize()

    # ---- value: state resident in HBM ------------------------------------
    for _ in range(args.warmup):
        solver.step(1)
    solver.profile(True)
    solver.profile_reset()
    launches0 = solver.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solver.step(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = solver.launch_count - launches0
    prof = solver.profile_read()
    solver.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the C ABI every step -----------------
    if slab is None:
        host = {f: torch.empty(solver._shape(f), dtype=torch.float64).pin_memory() for f in ("r", "v", "rho")}
        for f in host:
            solver.download_raw(f, host[f].data_ptr())
        h2d = d2h = sum(h.numel() * 8 for h in host.values())

        def e2e_step():
            for f in ("r", "v", "rho"):
                solver.upload_raw(f, host[f].data_ptr())
            solver.step(1)
            for f in ("r", "v", "rho"):
                solver.download_raw(f, host[f].data_ptr())
    else:
        # Slab mode: the owned fluid records (r, rho | v, m: 64 B per particle) travel
        # host -> device before and device -> host after every step.
        cap = int(slab.solver.mg_counts()[0] * 1.05) + 4096
        hostA = torch.empty((cap, 4), dtype=torch.float64).pin_memory()
        hostB = torch.empty((cap, 4), dtype=torch.float64).pin_memory()
        tdev = torch.device("cuda", local_rank)
        state = {"n": 0}

        def pull():
            with torch.cuda.stream(stream):
                rec, n_owned = slab._export(False)
                hostA[:n_owned].copy_(rec[:n_owned, 0:4], non_blocking=True)
                hostB[:n_owned].copy_(rec[:n_owned, 4:8], non_blocking=True)
            solver.synchronize()
            state["n"] = n_owned

        def push():
            k = state["n"]
            with torch.cuda.stream(stream):
                dA = hostA[:k].to(tdev, non_blocking=True)
                dB = hostB[:k].to(tdev, non_blocking=True)
                solver.mg_import(k, 0, dA.data_ptr(), dB.data_ptr())
                slab.gid = slab.gid[:k]
                slab._keep = [dA, dB]

        pull()
        h2d = d2h = int(state["n"]) * 64

        def e2e_step():
            push()
            solver.step(1)
            pull()

    # Between output frames the reference's time loop reads nothing but what the
    # next step needs (wcsph.cpp:170-193): the e2e loop publishes the state only.
    solver.set_outputs(0)
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, wall_ms)  # host-side packing counts too
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * n * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the kernel-sum pass ---------------------------------
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = solver.measure_fp64_peak()
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    total_kernel_ms = sum(v[1] for v in prof.values())
    rhs_name = next((k for k in prof if "k_rhs" in k), top[0][0] if top else None)
    roof = None
    if rhs_name:
        cnt, tot = prof[rhs_name]
        avg_s = tot / cnt * 1e-3
        ach = alg_bytes_rhs(dim) * n / avg_s / 1e9
        flops = alg_flops_rhs(dim) * (case.n_fluid if slab is None else case.meta["n_fluid_global"] / world) / avg_s / 1e12
        roof = {
            "bound": "hbm", "kernel": rhs_name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic(dim, n)[0], "traffic_unit": "bytes per launch (DRAM read + write)",
            "traffic_source": ncu_traffic(dim, n)[1],
            "peak_source": peak_src, "alg_bytes_per_particle": alg_bytes_rhs(dim), "launches": cnt, "avg_ms": avg_s * 1e3,
            "share_of_step": tot / total_kernel_ms if total_kernel_ms else None,
            "fp64": {"achieved_tflops": flops, "peak_tflops": fp64_peak, "frac": flops / fp64_peak if fp64_peak else None,
                     "alg_flops_per_particle": alg_flops_rhs(dim), "peak_source": "measured DFMA loop (titgpu_measure_fp64_peak)"},
            "step": {"alg_bytes_per_update": alg_bytes_step(dim), "achieved_gbs": alg_bytes_step(dim) * value / world / 1e9,
                     "frac": alg_bytes_step(dim) * value / world / 1e9 / hbm_peak},
        }

    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1)
        cpu, _, _, _ = cpu_baseline_sample(dim, args.cpu_budget)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (w["label"] if not args.n_col else f"{dim}D dam break, n_col={n_col}") + ("" if world == 1 else f", per GPU; {world} GPUs weak-scaled along z"), "particles_per_gpu": int(n), "n_fluid": case.n_fluid if slab is None else case.meta["n_fluid_global"], "n_fixed": case.n_fixed if slab is None else case.meta["n_fixed_global"],
                   "integrator": "ssprk3", "kernel": "SixthOrderWendland", "eos": "tait", "parallelism": "single GPU" if world == 1 else f"{world} slabs along z (tank {world}x deeper), one rank per GPU, ghost-layer exchange over NCCL send/recv before every neighbour search, dt all-reduce per step",
                   "outputs": "value: all 17 fields of all particles published after the last timed step (reference semantics); e2e: state only (r, v, rho) every step",
                   "l2_policy": "inputs larger than L2 (state arrays of %d MB)" % (int(n) * (2 * dim + 2) * 8 // 2**20)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in top[:12]},
    }
    emit(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


class QuietStdout:
    """Everything written to fd 1 while the benchmark runs (NCCL's version banner,
    library chatter) goes to stderr: stdout carries exactly the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit(line):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n-col", type=int, default=0, help="override the lattice resolution (parity / debugging runs)")
    ap.add_argument("--ref-n-col", type=int, default=0, help="sample resolution of the CPU reference arm")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global OUT
    with QuietStdout() as q:
        OUT = q
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_ours(args, rank, local_rank, world)
    OUT = None


if __name__ == "__main__":
    main()
