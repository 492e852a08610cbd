"""Builds matrixextra_b200/csrc/libmxgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
Every .cu is compiled to an object in parallel (the product kernels are heavily templated), then linked."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "build")
SOURCES = ["capi.cu", "layout.cu", "pipeline.cu", "spmm.cu", "spmv.cu", "transpose.cu", "synth.cu", "rowops.cu", "hoststage.cu", "residency.cu", "tmaprobe.cu"]
HEADERS = ["mxg_internal.cuh", os.path.join("..", "..", "include", "mxgpu.h")]
LIB = os.path.join(CSRC, "libmxgpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return _newest_header() > t or any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = find_nvcc()
    extra = os.environ.get("MXG_NVCC_EXTRA", "").split()  # development: e.g. -DSPMM_MINB=12
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc's host compiler is the system g++
    env.pop("CC", None)
    env.pop("CXX", None)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = _newest_header()

    def compile_one(src: str):
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        src_path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(src_path)):
            return obj, None
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-ccbin", "/usr/bin/g++", "-c", src, "-o", obj]
        res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    objs = []
    for obj, res in results:
        if res is not None:
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed building " + obj)
            if verbose:
                sys.stderr.write(res.stderr)
        objs.append(obj)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
    res = subprocess.run(link, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libmxgpu.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
