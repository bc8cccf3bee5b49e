"""In-tree build of libspano.so (hand-written CUDA for sm_100a + the C ABI).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
Usage: python -m simplepanorama_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libspano.so")

CU_SOURCES = ["warp_kernels.cu", "mask_kernels.cu", "blend_kernels.cu", "disk_kernels.cu", "resize_kernels.cu", "dist_kernels.cu", "equalize_kernels.cu", "capi.cu"]
CPP_SOURCES = ["projector_host.cpp"]
HEADERS = ["spano_internal.h", os.path.join("..", "..", "include", "spano.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
    # host code inside the .cu files (e.g. the disk_reproj geometry that decides integer tile corners) must round like the
    # reference's scalar code: no FMA contraction on FMA-default hosts either
    "-Xcompiler", "-ffp-contract=off",
]
# host geometry must round like OpenCV's SSE3-baseline build: no FMA contraction
CXX_FLAGS = ["-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-pthread"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _cuda_include() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], verbose: bool, log_name: str | None = None) -> None:
    p = subprocess.run(cmd, capture_output=True, text=True)
    if log_name:
        with open(os.path.join(OBJ, log_name), "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(p.stdout + p.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    hdrs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    jobs = []
    objs = []
    for s in CU_SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append(([_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj], s + ".log"))
    for s in CPP_SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append((["g++"] + CXX_FLAGS + ["-I", _cuda_include(), "-c", src, "-o", obj], s + ".log"))
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(lambda j: _run(j[0], verbose, j[1]), jobs))
    if jobs or force or _stale(LIB, objs):
        _run([_nvcc(), "-shared", "-o", LIB] + objs + ["-Xcompiler", "-pthread", "-cudart", "static"], verbose, "link.log")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
