"""CPU-side checks of bench.py's contract: the reference arm prints exactly one
JSON line with the agreed keys, and the roofline helpers use SURVEY.md §8(d)'s
per-unit figures. (The GPU arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_and_flops():
    assert bench.alg_bytes_rhs(2) == 96 and bench.alg_bytes_rhs(3) == 128
    assert bench.alg_bytes_step(2) == 1280 and bench.alg_bytes_step(3) == 1776
    assert bench.alg_flops_rhs(3) == 256 * 90


def test_ncu_traffic_comes_from_the_committed_capture():
    t, src = bench.ncu_traffic(3, 1000)
    assert t is not None and t > 128 * 1000  # DRAM traffic is above the algorithmic bytes
    assert "profiles/" in src and os.path.exists(os.path.join(ROOT, src.split(":")[0]))
    t2, src2 = bench.ncu_traffic(2, 1000)
    assert t2 > 96 * 1000 and os.path.exists(os.path.join(ROOT, src2.split(":")[0]))
    pipes = bench.ncu_pipes(3)
    assert 0 < pipes["fp64_pct"] < 100 and 0 < pipes["fma_fp32_pct"] < 100 and 0 < pipes["issue_slots_pct"] <= 100


def test_both_arms_share_one_config():
    """`same_config`: the reference arm describes the workload with the GPU arm's own function."""
    import argparse

    args = argparse.Namespace(n_col=0)
    w = bench.WORKLOADS["c3"]
    from titsolver_b200 import cases

    nf, nx = cases.dam_break_3d_counts(w["n_col"])
    cfg = bench.bench_config(w, args, 3, w["n_col"], 1, nf, nx, False)
    assert cfg["workload"] == w["label"] and cfg["n_fluid"] == 9941940 and cfg["n_fixed"] == 1803710 and cfg["particles_per_gpu"] == nf + nx
    cfg8 = bench.bench_config(w, args, 3, w["n_col"], 8, *cases.dam_break_3d_counts(w["n_col"], (5.366, 4.0, 8.0)), False)
    assert "8 GPUs weak-scaled" in cfg8["workload"] and cfg8["n_fluid"] == 79944894


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = {**os.environ, "RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
