"""Host-side check of the index maps the kernels rely on (no GPU): the y-tiled cell order of the 3-D
search grid (csrc/engine.cuh: col_base / cell_flat / cell_unflat, `__host__ __device__`) and the
particle-to-(chunk, lane) map of the thread-scan kernels (ScanMap). tests/cpp/cell_order_check.cu
includes the product header and is compiled with nvcc for the host."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")


def test_cell_order_and_scan_map_are_bijections(tmp_path):
    exe = str(tmp_path / "cell_order_check")
    r = subprocess.run([NVCC, "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-ccbin", "/usr/bin/g++",
                        "-I", os.path.join(ROOT, "titsolver_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-DTIT_D=3", "-DTIT_K=4",
                        "-o", exe, os.path.join(ROOT, "tests", "cpp", "cell_order_check.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
