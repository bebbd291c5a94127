"""CPU: the C-ABI library loads and exports every symbol include/pointops_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def header_symbols():
    with open(os.path.join(REPO, "include", "pointops_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pops_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pointops_b200.h but not exported"
    # and the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == syms


def test_library_identifies_itself():
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    assert lib.pops_abi_version() == 1
    assert b"sm_100a" in lib.pops_build_info()
    assert lib.pops_launch_count() >= 0


def test_version_predicates_match_reference_contract():
    # knn.cu:292-303: v0 always, v1 D<=32, v2 D<=8 & K<=32, v3 D<=8 & K<=4
    from pytorch3d_pointops_b200 import _C

    assert _C.knn_check_version(0, 500, 500)
    assert _C.knn_check_version(1, 32, 100) and not _C.knn_check_version(1, 33, 1)
    assert _C.knn_check_version(2, 8, 32) and not _C.knn_check_version(2, 8, 33)
    assert _C.knn_check_version(3, 3, 4) and not _C.knn_check_version(3, 3, 5)


def test_sass_is_blackwell_native():
    """The shipped cubin is sm_100a and uses TMA bulk copies + packed FP32 FMA."""
    import shutil
    import subprocess

    from pytorch3d_pointops_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out, "TMA bulk copy (cp.async.bulk) missing from SASS"
    assert "FFMA2" in out, "packed FP32 FMA missing from SASS"
    # tensor-core KNN (csrc/knn_tc.cu): tcgen05.mma / tcgen05.ld / TMA tensor loads
    assert re.search(r"\bUTC[A-Z]*MMA\b", out), "tcgen05.mma missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"
    assert "UTMALDG" in out, "cp.async.bulk.tensor missing from SASS"


def test_exact_kernels_have_no_fused_multiply_add():
    """Kernels that compute the reference's unfused distance must contain no FFMA/FFMA2 except
    where a fused filter is intended (the D=3 L2 scan) -- guards against compiler contraction
    (ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false)."""
    import shutil
    import subprocess

    from pytorch3d_pointops_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
    fn, bad = None, []
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif fn and re.search(r"\bFFMA2?\b", line):
            # knn_scan_kernel<DT,NORM,EXP,...>: EXP=false variants scan with the exact distance;
            # knn_flush_one (every variant) re-evaluates candidates exactly
            exact = re.search(r"knn_scan_kernelILi\d+ELi\d+ELb0", fn) or "knn_flush_one" in fn \
                or "knn_generic_kernel" in fn or "ball_query_generic" in fn or "fps_" in fn \
                or "knn_backward" in fn or "prune_flush_one" in fn or "seed_bound" in fn \
                or "knn_exact_rows_kernel" in fn
            # (knn_tc_rerank_kernel is not listed: its sqrtf() for the error bound lowers to FFMA;
            #  its distance uses __fsub_rn/__fmul_rn/__fadd_rn, which are never contracted)
            if exact:
                bad.append(fn)
    assert not bad, sorted(set(bad))


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly: the product has no CPU path."""
    import torch

    from pytorch3d_pointops_b200.functions import ball_query, knn_points, sample_farthest_points

    p = torch.rand(1, 8, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        knn_points(p, p, K=2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ball_query(p, p, K=2, radius=0.5)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        sample_farthest_points(p, K=2)


def test_product_never_imports_oracle():
    """Nothing under the package (or the alias package) may reference oracle/."""
    for root in ("pytorch3d_pointops_b200", "pytorch3d_pointops"):
        for dirpath, _, files in os.walk(os.path.join(REPO, root)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    with open(os.path.join(dirpath, f)) as fh:
                        src = fh.read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), (dirpath, f)
                    assert "_C_ref" not in src, (dirpath, f)
