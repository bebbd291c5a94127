"""Compile the UNMODIFIED reference CPU sources into oracle/_ref/ (test infrastructure).

The reference (mikel-zhobro/pytorch3d_pointops) ships its CPU path as a handful of
self-contained translation units under pytorch3d_pointops/csrc (ext.cpp + <op>_cpu.cpp).
This script runs g++ on those files *where they lie* under /root/reference -- no source is
copied, the reference's own setup.py is not run -- and writes one CPython extension module
`oracle/_ref/_C_ref.<abi>.so` exporting the reference's 7 CPU entry points
(ext.cpp:15-27 without WITH_CUDA).  Flags mirror the reference build: -O2 -std=c++17,
no -march / -mfma, so the objects contain no fused multiply-adds.

oracle/_ref/ is git-ignored but travels to the GPU box with gpurun (same image, same torch).
Only tests/, tests/golden/make_golden.py, __graft_entry__ and bench.py's reference /
cpu_baseline legs may load it.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("POINTOPS_REFERENCE_ROOT", "/root/reference")
CSRC = os.path.join(REF_ROOT, "pytorch3d_pointops", "csrc")
OUT_DIR = os.path.join(HERE, "_ref")
MODULE = "_C_ref"

SOURCES = [
    "ext.cpp",
    "knn/knn_cpu.cpp",
    "ball_query/ball_query_cpu.cpp",
    "sample_farthest_points/sample_farthest_points_cpu.cpp",
    "packed_to_padded_tensor/packed_to_padded_tensor_cpu.cpp",
    "sample_pdf/sample_pdf_cpu.cpp",
]


def ref_so_path() -> str:
    return os.path.join(OUT_DIR, MODULE + sysconfig.get_config_var("EXT_SUFFIX"))


def available() -> bool:
    return os.path.isfile(ref_so_path())


def build(force: bool = False, verbose: bool = True) -> str | None:
    """Returns the path of the built module, or None when the reference tree is absent."""
    out = ref_so_path()
    if os.path.isfile(out) and not force:
        return out
    if not os.path.isdir(CSRC):
        if verbose:
            print(f"[build_ref] {CSRC} not present; skipping (prebuilt file expected)")
        return None
    from torch.utils import cpp_extension  # only needed to locate headers / libs

    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    inc = [f"-I{CSRC}"] + [f"-isystem{p}" for p in cpp_extension.include_paths()]
    inc.append(f"-isystem{sysconfig.get_paths()['include']}")
    cxx = os.environ.get("CXX", "g++")
    common = [
        cxx, "-O2", "-std=c++17", "-fPIC", "-w",
        f"-DTORCH_EXTENSION_NAME={MODULE}", "-DTORCH_API_INCLUDE_EXTENSION_H",
    ]

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, "obj", src.replace("/", "_") + ".o")
        cmd = common + inc + ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    libdirs = cpp_extension.library_paths()
    link = [cxx, "-shared", "-o", out] + objs
    for d in libdirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    subprocess.run(link, check=True)
    if verbose:
        print(f"[build_ref] built {out}")
    return out


def load():
    """Import oracle/_ref/_C_ref as a module (torch must be imported first)."""
    import importlib.util

    import torch  # noqa: F401  (registers the libtorch symbols the module links against)

    path = ref_so_path()
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(MODULE, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    sys.exit(0 if (p or not os.path.isdir(CSRC)) else 1)
