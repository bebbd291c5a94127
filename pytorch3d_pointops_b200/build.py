"""Build libpointops_b200.so in-tree with nvcc for sm_100a.

    python -m pytorch3d_pointops_b200.build [--force] [--verbose]

The library is plain CUDA behind a C ABI (include/pointops_b200.h); it does not link against
torch, so it builds in seconds per file and cross-compiles on a machine without a GPU.  The
`.so` is written next to the package (pytorch3d_pointops_b200/lib/) so that it travels with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libpointops_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # never contract a*b+c: the reference arithmetic is unfused and ptxas was seen
    # fusing mul.rn.f32x2 + add.rn.f32x2 into FFMA2.  Explicit fmaf()/__ffma2_rn() are unaffected.
    "-fmad=false",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(cand):
        raise RuntimeError("nvcc not found; libpointops_b200.so cannot be built")
    return cand


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/pointops_b200.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(LIB_DIR, "build.stamp")
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()
    logs = {}

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    link = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as fh:
        for src in sorted(logs):
            fh.write(f"==== {src}\n{logs[src]}\n")
    with open(os.path.join(LIB_DIR, "build.stamp"), "w") as fh:
        fh.write(_digest())
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
