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

REBALANCE_SHIFT = 30  # particle spacings a slab edge may move away from the initial equal-count cut
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


def ncu_record(dim, kernel):
    """The committed `ncu --set full` capture of the kernel-sum pass that ran (profiles/ncu_traffic.json,
    written by tools/ncu_traffic_update.py): the grouped sweep `k_rhs_grp` on large particle counts, the
    gather traversal `k_rhs` otherwise."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            db = json.load(f)
    except (OSError, ValueError):
        return None
    key = "k_rhs_grp" if kernel and "k_rhs_grp" in kernel else "k_rhs"
    return db.get(key, {}).get(f"{dim}d")


def ncu_pipes(dim, kernel=None):
    """Pipe utilisation of that launch: FP64 / FP32 (FMA) pipe and L1 data-pipe busy %, issue-slot utilisation."""
    rec = ncu_record(dim, kernel)
    return rec.get("pipes") if rec else None


def ncu_traffic(dim, n, kernel=None):
    """DRAM bytes (read + write) of one launch over n particles (bytes per particle measured at the
    capture's size, scaled to this launch); None when no capture exists for this kernel and dimension."""
    rec = ncu_record(dim, kernel)
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


class CpuReference:
    """The reference's CPU algorithm (oracle/, -O3 OpenMP build, all host cores) timed on the
    benchmark workload.

    2-D workloads run as they are. The 3-D workloads (>= 10 M particles, minutes per step on a
    CPU) are timed through two bounded samples whose per-phase costs recombine to the workload's
    own mixture of fluid and wall particles:
      bulk sample   a wall-free block of fluid at the workload's spacing: seconds per FLUID
                    particle of every phase (search, pair sums, shifting ...);
      wall sample   the closed dam-break tank at n_col = 40: what is left of every phase after
                    the bulk part, per WALL particle (the boundary integrals of compute_gamma and
                    of the face terms scale with the wetted / meshed wall area).
    step time of the workload = sum over phases of a_phase n_fluid + b_phase n_fixed. (A single
    small tank mis-states the mixture: at n_col = 28 it holds 53 % wall particles, C3 15 %.)"""

    PHASES = ("search", "compute_gamma", "setup_boundary", "continuity_momentum", "update_lincomb_dt", "apply_shifts", "free_surface_correction")

    def __init__(self, dim, n_col, tank_z=1.0, bulk=(96, 64, 64), wall_n_col=40):
        import oracle_lib
        from titsolver_b200 import cases

        self.dim, self.n_col = dim, n_col
        self.threads = host_threads()

        def make(case):
            s = oracle_lib.OracleSolver(dim, fast=True)
            s.set_symmetric(os.environ.get("BENCH_CPU_GATHER") is None)  # the reference's loop structure: every pair once, both particles updated
            s.lib.orc_set_num_threads(self.threads)  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
            oracle_lib.load_case(s, case)
            s.initialize()
            return s

        if dim == 2:
            case = cases.dam_break_2d(n_col)
            self.direct = make(case)
            self.n_fluid, self.n_fixed = case.n_fluid, case.n_fixed
            self.sample = f"the workload itself ({case.n} particles)"
        else:
            self.direct = None
            self.n_fluid, self.n_fixed = cases.dam_break_3d_counts(n_col, tank=(5.366, 4.0, tank_z))
            cb, cw = cases.fluid_block_3d(*bulk, n_col=n_col), cases.dam_break_3d(wall_n_col)
            self.bulk, self.wall = make(cb), make(cw)
            self.nb, self.nwf, self.nwx = cb.n_fluid, cw.n_fluid, cw.n_fixed
            self.sample = (f"per-phase costs of two samples recombined to the workload's {self.n_fluid} fluid + {self.n_fixed} wall particles: a wall-free fluid block "
                           f"{bulk[0]}x{bulk[1]}x{bulk[2]} ({cb.n_fluid} particles: seconds per fluid particle) and the closed tank at n_col={wall_n_col} "
                           f"({cw.n_fluid} fluid + {cw.n_fixed} wall particles: the remainder per wall particle)")
        self.cores = int(oracle_lib.load(True).orc_num_threads())
        self.phase_s = {k: 0.0 for k in self.PHASES}

    @property
    def n(self):
        return self.n_fluid + self.n_fixed

    def step(self):
        """One SSPRK3 step of the workload: measured (2-D) or recombined from one step of each sample (3-D). Seconds."""
        if self.direct is not None:
            self.direct.phase_times()
            t0 = time.perf_counter()
            self.direct.step(1)
            el = time.perf_counter() - t0
            for k, v in self.direct.phase_times().items():
                self.phase_s[k] += v
            return el
        self.bulk.phase_times(); self.wall.phase_times()
        self.bulk.step(1); self.wall.step(1)
        pb, pw = self.bulk.phase_times(), self.wall.phase_times()
        total = 0.0
        for k in self.PHASES:
            a = pb[k] / self.nb
            b = max(pw[k] - a * self.nwf, 0.0) / self.nwx
            t = a * self.n_fluid + b * self.n_fixed
            self.phase_s[k] += t
            total += t
        return total


def cpu_baseline_sample(dim, n_col, budget_s=30.0, tank_z=1.0):
    """The cpu_baseline object of the GPU arm's line: `CpuReference` for about `budget_s` seconds."""
    ref = CpuReference(dim, n_col, tank_z)
    ref.step()  # warm-up (first touch, thread pool)
    for k in ref.phase_s:
        ref.phase_s[k] = 0.0
    steps, est, t0 = 0, 0.0, time.perf_counter()
    while True:
        est += ref.step()
        steps += 1
        if time.perf_counter() - t0 >= budget_s or steps >= 20:
            break
    return {"value": ref.n * steps / est, "unit": UNIT, "cores": ref.cores, "kind": "port", "seconds_per_step": est / steps,
            "phase_seconds_per_step": {k: round(v / steps, 4) for k, v in ref.phase_s.items()},
            "sample": f"{steps} SSPRK3 steps; {ref.sample}; oracle/liboracle_fast.so (-O3, OpenMP, symmetric block-coloured pair sums as in particle_mesh.hpp:165-241)"}


def bench_config(w, args, dim, n_col, world, n_fluid, n_fixed, strong_main):
    """The `config` object, the same for both arms."""
    n_job = n_fluid + n_fixed
    if world == 1:
        par = "single GPU"
    elif strong_main:
        par = f"{world} slabs along x of the fixed tank (strong scaling), one rank per GPU"
    else:
        par = f"{world} slabs along z (tank {world}x deeper: weak scaling), one rank per GPU"
    if world > 1:
        par += "; inside every step: migration + halo set, 4 ghost refreshes, {N, phi} and shifted-record exchanges (ncclSend/ncclRecv on the context's stream), dt all-reduce"
    label = w["label"] if not args.n_col else f"{dim}D dam break, n_col={n_col}"
    if world > 1 and not strong_main:
        label += f", per GPU; {world} GPUs weak-scaled along z"
    return {"workload": label, "particles_per_gpu": int(n_job / world), "n_fluid": int(n_fluid), "n_fixed": int(n_fixed),
            "integrator": "ssprk3", "kernel": "SixthOrderWendland", "eos": "tait", "parallelism": par,
            "outputs": "value: all 17 fields of all particles published after the last timed step (reference semantics); e2e: state only (r, v, rho) every step",
            "l2_policy": "inputs larger than L2 (state arrays of %d MB)" % (int(n_job / world) * (2 * dim + 2) * 8 // 2**20)}


def run_reference(args, rank, world):
    """The reference arm: the CPU algorithm alone, on the GPU arm's workload and config."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    dim = w["dim"]
    n_col = args.n_col or w["n_col"]
    strong_main = args.scaling == "strong"
    tank_z = float(world) if (dim == 3 and world > 1 and not strong_main) else 1.0
    ref = CpuReference(dim, n_col, tank_z)
    for _ in range(args.warmup):
        ref.step()
    for k in ref.phase_s:
        ref.phase_s[k] = 0.0
    t0 = time.perf_counter()
    est = sum(ref.step() for _ in range(args.steps))
    wall = time.perf_counter() - t0
    val = ref.n * args.steps / est
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": est / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if strong_main else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(w, args, dim, n_col, world, ref.n_fluid, ref.n_fixed, strong_main),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": "port", "seconds_per_step": est / args.steps, "measured_wall_seconds": wall,
                         "phase_seconds_per_step": {k: round(v / args.steps, 4) for k, v in ref.phase_s.items()},
                         "sample": f"{args.steps} SSPRK3 steps; {ref.sample}; oracle/liboracle_fast.so (-O3, OpenMP, symmetric block-coloured pair sums as in particle_mesh.hpp:165-241)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference (OpenMP, all host cores); the reference itself (C++26, oneTBB) cannot be built in this image",
    }
    emit(json.dumps(out))


def strong_reference():
    """The committed single-GPU rate of the strong-scaling workload (C5), measured with
    `bench.py --scaling strong --workload c5 --gpus 1` (profiles/strong_c5_1gpu.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "strong_c5_1gpu.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


class Runner:
    """One workload on this rank's GPU: a whole case (world == 1 or `whole`) or one slab."""

    def __init__(self, args, dim, n_col, rank, local_rank, world, mode):
        import torch

        import titsolver_b200 as tb

        self.torch, self.world, self.rank, self.dim, self.mode = torch, world, rank, dim, mode
        self.slab = None
        if world > 1:
            if dim != 3:
                raise SystemExit("bench.py: the multi-GPU workload is the 3-D dam break (use --workload c3/c4/c5)")
            from titsolver_b200 import cases
            from titsolver_b200.slab import SlabSolver

            # strong scaling: the wall pieces reach REBALANCE_SHIFT spacings beyond the usual margin, so that
            # the slab edges can follow the measured cost (SlabSolver.rebalance)
            case, edges, axis = cases.dam_break_3d_slab(n_col, world, rank, mode=mode, halo_cells=18 + (REBALANCE_SHIFT + 2 if mode == "strong_x" else 0))
            self.slab = SlabSolver(case, rank, world, axis=axis, edges=edges, device=local_rank, local=True)
            self.extent = (case.dr, case.dr * 2 * n_col)  # the fluid column along x at t = 0
            self.dr = case.dr
            self.solver = self.slab.solver
            self.n_fluid, self.n_fixed = case.meta["n_fluid_global"], case.meta["n_fixed_global"]
        else:
            case = make_case(dim, n_col)
            self.solver = tb.Solver(dim, device=local_rank)
            tb.load_case(self.solver, case)
            self.n_fluid, self.n_fixed = case.n_fluid, case.n_fixed
        self.n_total = self.n_fluid + self.n_fixed  # particles of the whole job
        self.name = case.meta.get("name", "")
        del case
        self.solver.initialize()
        self.stream = torch.cuda.ExternalStream(self.solver.stream, device=torch.device("cuda", local_rank))
        self.local_rank = local_rank

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        self.torch.cuda.synchronize()
        self.solver.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(self, steps, warmup, profile=True):
        """W untimed + K timed steps with the state resident in HBM; device time, max over ranks."""
        torch, solver = self.torch, self.solver
        for _ in range(warmup):
            solver.step(1)
        if profile:
            solver.profile(True)
            solver.profile_reset()
        launches0 = solver.launch_count
        sampler = ClockSampler(self.local_rank)
        self.barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        solver.step(steps)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = solver.launch_count - launches0
        prof = solver.profile_read() if profile else {}
        if profile:
            solver.profile(False)
        self.rank_info = {"rank": self.rank, "ms_per_step": ms / steps, "kernel_ms_per_step": sum(v[1] for v in prof.values()) / steps if prof else None,
                          "counts": self.solver.mg_counts(), "exchanges_migrated": self.solver.mg_stats()}
        ms = self.max_over_ranks(ms)
        return ms, clocks, launches, prof

    def rebalance(self, rounds=2, steps=5):
        """Strong scaling: measure the kernel time per rank over `steps` steps, move the slab edges so
        that every rank carries the same cost, let the next steps migrate the particles; `rounds` times.
        (The end slabs carry the end walls and, the last one, the dry part of the tank.)"""
        if self.slab is None or self.mode != "strong_x":
            return None
        solver, hist = self.solver, []
        for _ in range(rounds):  # (same call pattern as the timed region: the last step of a call publishes all fields)
            solver.profile(True)
            solver.profile_reset()
            solver.step(steps)
            cost = sum(v[1] for v in solver.profile_read().values()) / steps
            solver.profile(False)
            edges = self.slab.rebalance(cost, self.extent, max_shift=REBALANCE_SHIFT * self.dr)
            hist.append({"cost_ms": cost, "edges_dr": [round(e / self.dr, 2) for e in edges[1:-1]]})
            solver.step(1)  # migration to the new slabs
        return hist

    def gather_rank_info(self):
        """Per-rank step time, kernel-time sum and particle counts (owned, ghosts, walls) on rank 0."""
        if self.world == 1:
            return [self.rank_info]
        import torch.distributed as dist

        out = [None] * self.world
        dist.all_gather_object(out, self.rank_info)
        return out

    def e2e_steps(self, steps, warmup):
        """K steps through the C ABI with HOST buffers: every step uploads the state from
        pinned memory, steps, and downloads it again. Returns (ms, h2d bytes, d2h bytes)."""
        torch, solver = self.torch, self.solver
        if self.slab is None:
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
            # Slab mode: the owned fluid records (r, rho | v, m: 64 B per particle) and their global
            # ids travel host -> device before and device -> host after every step (titgpu_mg_upload_owned /
            # titgpu_mg_download_owned); the ghosts are fetched from the neighbours inside the step.
            cap = int(solver.mg_counts()[0] * 1.05) + 4096
            hostA = torch.empty((cap, 4), dtype=torch.float64).pin_memory()
            hostB = torch.empty((cap, 4), dtype=torch.float64).pin_memory()
            hostG = torch.empty((cap,), dtype=torch.int64).pin_memory()
            state = {"n": solver.mg_download_owned(hostA.data_ptr(), hostB.data_ptr(), hostG.data_ptr(), cap=cap)}
            h2d = d2h = int(state["n"]) * 72

            def e2e_step():
                solver.mg_upload_owned(state["n"], hostA.data_ptr(), hostB.data_ptr(), hostG.data_ptr())
                solver.step(1)
                state["n"] = solver.mg_download_owned(hostA.data_ptr(), hostB.data_ptr(), hostG.data_ptr(), cap=cap)

        # Between output frames the reference's time loop reads nothing but what the
        # next step needs (wcsph.cpp:170-193): the e2e loop publishes the state only.
        solver.set_outputs(0)
        for _ in range(min(warmup, 3)):
            e2e_step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            e2e_step()
        e1.record(self.stream)
        self.barrier()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)  # host-side work counts too
        solver.set_outputs(2)
        return self.max_over_ranks(ms), h2d, d2h

    def close(self):
        self.solver.close()
        self.slab = None


def run_ours(args, rank, local_rank, world):
    import torch

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
    strong_main = args.scaling == "strong"
    mode = "strong_x" if strong_main else "weak_z"
    run = Runner(args, dim, n_col, rank, local_rank, world, mode)
    solver = run.solver
    balance = run.rebalance(steps=args.steps) if not args.no_rebalance else None
    n_job, n_fluid, n_fixed = run.n_total, run.n_fluid, run.n_fixed  # particles of the whole job
    n = n_job / world  # per GPU (wall particles of the halos are not counted twice)

    ms, clocks, launches, prof = run.timed_steps(args.steps, args.warmup)
    value = n_job * args.steps / (ms * 1e-3)
    ranks = run.gather_rank_info()
    e2e_ms, h2d, d2h = run.e2e_steps(args.steps, args.warmup)
    e2e_value = n_job * args.steps / (e2e_ms * 1e-3)

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
        flops = alg_flops_rhs(dim) * (n_fluid / world) / avg_s / 1e12
        traffic, traffic_src = ncu_traffic(dim, n, rhs_name)
        roof = {
            # The kernel-sum pass is bound by the FP64 pipe / instruction issue, not by HBM: its
            # algorithmic intensity is ~180 flop per compulsory byte. The HBM
            # figure BASELINE.json asks for is reported beside it.
            "bound": "fp64", "kernel": rhs_name, "achieved": flops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops / fp64_peak if fp64_peak else None,
            "peak_source": "measured DFMA loop on all SMs (titgpu_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
            "alg_flops_per_particle": alg_flops_rhs(dim), "launches": cnt, "avg_ms": avg_s * 1e3,
            "share_of_step": tot / total_kernel_ms if total_kernel_ms else None,
            "traffic": traffic, "traffic_unit": "bytes per launch (DRAM read + write)", "traffic_source": traffic_src,
            "hbm": {"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "alg_bytes_per_particle": alg_bytes_rhs(dim), "peak_source": peak_src,
                    "dram_frac": (traffic / avg_s / 1e9 / hbm_peak) if traffic else None},
            "pipes": ncu_pipes(dim, rhs_name),
            "step": {"alg_bytes_per_update": alg_bytes_step(dim), "achieved_gbs": alg_bytes_step(dim) * value / world / 1e9,
                     "frac": alg_bytes_step(dim) * value / world / 1e9 / hbm_peak},
        }

    # ---- strong scaling beside the weak line (N > 1) ----------------------
    strong = None
    if world > 1 and not strong_main and not args.no_strong:
        run.close()
        del run, solver
        sw = WORKLOADS[args.strong_workload]
        srun = Runner(args, 3, args.strong_n_col or sw["n_col"], rank, local_rank, world, "strong_x")
        s_steps = max(2, min(args.steps, 5))
        s_balance = srun.rebalance(steps=s_steps) if not args.no_rebalance else None
        s_ms, _, _, s_prof = srun.timed_steps(s_steps, 3, profile=True)
        s_value = srun.n_total * s_steps / (s_ms * 1e-3)
        s_ranks = srun.gather_rank_info()
        ref = strong_reference()
        same = bool(ref) and ref.get("n_total") == srun.n_total
        strong = {"workload": sw["label"] if not args.strong_n_col else f"3D dam break, n_col={args.strong_n_col}", "n_total": srun.n_total, "value": s_value, "unit": UNIT, "steps": s_steps, "ms_per_step": s_ms / s_steps,
                  "decomposition": f"{world} slabs of equally many lattice planes along x (fixed tank), halo 2R + dr_wall + dr",
                  "one_gpu_value": ref["value"] if same else None, "one_gpu_source": "profiles/strong_c5_1gpu.json" if same else None,
                  "efficiency": (s_value / (world * ref["value"])) if same else None, "ranks": s_ranks, "rebalance": s_balance,
                  "kernels_ms_per_step_rank0": {k: round(v[1] / s_steps, 3) for k, v in sorted(s_prof.items(), key=lambda kv: -kv[1][1])}}
        srun.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1)
        cpu = cpu_baseline_sample(dim, n_col, args.cpu_budget)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong_main else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(w, args, dim, n_col, world, n_fluid, n_fixed, strong_main),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in top[:12]},
    }
    if world > 1:
        out["ranks"] = ranks
    if balance is not None:
        out["rebalance"] = balance
    if strong is not None:
        out["strong"] = strong
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
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload per GPU (tank N x deeper along z); strong = the workload itself cut into N slabs along x")
    ap.add_argument("--no-strong", action="store_true", help="N > 1, weak: skip the extra strong-scaling measurement (`strong` object)")
    ap.add_argument("--no-rebalance", action="store_true", help="strong scaling: keep the equal-count slabs (no cost-based rebalancing)")
    ap.add_argument("--strong-workload", default="c5", choices=["c3", "c4", "c5"])
    ap.add_argument("--strong-n-col", type=int, default=0)
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
