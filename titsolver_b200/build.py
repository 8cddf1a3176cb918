"""Build libtitgpu.so (hand-written sm_100a CUDA + the C ABI) in-tree.

    python -m titsolver_b200.build [--dims 2,3] [--kernels 0,1,2,3,4,5] [-j N]

nvcc cross-compiles without a GPU. The shared library is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# TITGPU_VARIANT / TITGPU_DEFINES build a tuning variant next to the product library
# (libtitgpu_<variant>.so, selected at run time with TITGPU_LIB).
VARIANT = os.environ.get("TITGPU_VARIANT", "")
DEFINES = [f"-D{d}" for d in os.environ.get("TITGPU_DEFINES", "").split() if d]
OBJ = os.path.join(HERE, "_build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libtitgpu" + ("_" + VARIANT if VARIANT else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ccbin", "/usr/bin/g++"]
HEADERS = ["common.cuh", "context.h", "engine.cuh", "mg.cuh", "tile.cuh", "mg_transport.h", "sph_kernel.cuh", "kernels_gen.cuh", "../../include/titgpu.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    return r.stdout + r.stderr


def build(dims=(2, 3), kernels=(0, 1, 2, 3, 4, 5), jobs=None, verbose=False, ptxas_v=False):
    os.makedirs(OBJ, exist_ok=True)
    gen = os.path.join(CSRC, "kernels_gen.cuh")
    gen_src = os.path.join(os.path.dirname(HERE), "tools", "gen_kernels.py")
    if _newer(gen, [gen_src]):
        _run([sys.executable, gen_src, "--product", gen])
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    tasks = []
    api_o = os.path.join(OBJ, "titgpu_api.o")
    api_src = os.path.join(CSRC, "titgpu_api.cu")
    if _newer(api_o, [api_src] + hdrs):
        tasks.append([NVCC, *ARCH, *FLAGS, "-c", api_src, "-o", api_o])
    objs = [api_o]
    mgt_o = os.path.join(OBJ, "mg_transport.o")
    mgt_src = os.path.join(CSRC, "mg_transport.cu")
    if _newer(mgt_o, [mgt_src, os.path.join(CSRC, "mg_transport.h")]):
        tasks.append([NVCC, *ARCH, *FLAGS, "-c", mgt_src, "-o", mgt_o])
    objs.append(mgt_o)
    inst = os.path.join(CSRC, "inst.cu")
    extra = ["-Xptxas", "-v"] if ptxas_v else []
    for d in dims:
        for k in kernels:
            o = os.path.join(OBJ, f"inst_{d}_{k}.o")
            objs.append(o)
            if _newer(o, [inst] + hdrs) or ptxas_v:
                tasks.append([NVCC, *ARCH, *FLAGS, *extra, *DEFINES, f"-DTIT_D={d}", f"-DTIT_K={k}", "-c", inst, "-o", o])
    outs = []
    if tasks:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
            outs = list(ex.map(_run, tasks))
    if tasks or _newer(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-ccbin", "/usr/bin/g++"])
    if verbose or ptxas_v:
        for o in outs:
            if o.strip():
                print(o)
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="2,3")
    ap.add_argument("--kernels", default="0,1,2,3,4,5")
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--ptxas-v", action="store_true")
    a = ap.parse_args()
    lib = build(tuple(int(x) for x in a.dims.split(",")), tuple(int(x) for x in a.kernels.split(",")), a.j, a.v, a.ptxas_v)
    print(lib)


if __name__ == "__main__":
    main()
