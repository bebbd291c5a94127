"""ctypes binding of libpointops_b200.so (the C ABI declared in include/pointops_b200.h).

There is no CPU or pure-PyTorch fallback: if the shared library cannot be loaded (or built with
nvcc when missing) importing any op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

from . import build as _build

_LIB = None

# name -> (restype, argtypes); mirrors include/pointops_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "pops_abi_version": (c_int, []),
    "pops_build_info": (c_char_p, []),
    "pops_last_error": (c_char_p, []),
    "pops_launch_count": (c_int64, []),
    "pops_set_option": (None, [c_char_p, c_int]),
    "pops_knn_debug_stats": (c_int, [_P]),
    "pops_profile_enable": (None, [c_int]),
    "pops_profile_reset": (None, []),
    "pops_profile_read": (c_int, [c_char_p, _P, _P]),
    "pops_fp32_peak_probe": (ctypes.c_double, [c_int, _P]),
    "pops_knn_workspace_bytes": (c_size_t, [c_int64] * 5 + [c_int]),
    "pops_knn_points_idx": (c_int, [_P, _P, _P, _P] + [c_int64] * 5 + [c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "pops_knn_points_prepare": (c_int, [_P, _P, _P, _P] + [c_int64] * 5 + [c_int, _P, c_size_t, _P]),
    "pops_knn_points_idx_range": (c_int, [_P, _P, _P, _P] + [c_int64] * 5 + [c_int, c_int] + [c_int64] * 2
                                  + [_P, _P, _P, c_size_t, _P]),
    "pops_knn_pair_workspace_bytes": (c_size_t, [c_int64] * 5 + [c_int]),
    "pops_knn_points_idx_pair": (c_int, [_P, _P, _P, _P] + [c_int64] * 5 + [c_int] + [_P] * 5 + [c_size_t, _P]),
    "pops_knn_check_version": (c_int, [c_int, c_int64, c_int64]),
    "pops_knn_points_backward": (c_int, [_P] * 6 + [c_int64] * 5 + [c_int, _P, _P, _P]),
    "pops_knn_backward_workspace_bytes": (c_size_t, [c_int64] * 3),
    "pops_knn_points_backward_ws": (c_int, [_P] * 6 + [c_int64] * 5 + [c_int, _P, _P, _P, c_size_t, _P]),
    "pops_ball_query_workspace_bytes": (c_size_t, [c_int64] * 5),
    "pops_ball_query": (c_int, [_P] * 4 + [c_int64] * 5 + [c_float, _P, _P, _P, c_size_t, _P]),
    "pops_fps_workspace_bytes": (c_size_t, [c_int64] * 4),
    "pops_sample_farthest_points": (c_int, [_P] * 4 + [c_int64] * 4 + [_P, _P, c_size_t, _P]),
    "pops_packed_to_padded": (c_int, [_P, _P] + [c_int64] * 4 + [_P, _P]),
    "pops_padded_to_packed": (c_int, [_P, _P] + [c_int64] * 4 + [_P, _P]),
    "pops_sample_pdf": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_float, _P]),
    "pops_point_covariances": (c_int, [_P, _P, _P] + [c_int64] * 5 + [_P, _P, _P]),
    "pops_gather": (c_int, [_P, _P, _P] + [c_int64] * 5 + [c_int, _P, _P, _P]),
    "pops_gather_backward": (c_int, [_P, _P, _P] + [c_int64] * 5 + [c_int, _P, _P]),
    "pops_chamfer_forward": (c_int, [_P] * 5 + [c_int64] * 3 + [c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P]),
    "pops_chamfer_backward": (c_int, [_P] * 6 + [c_int64] * 4 + [c_int, c_int, _P, _P, _P, c_int, c_int]
                              + [_P] * 7 + [c_int, c_int, c_float, _P]),
}


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing) and type the C ABI."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.isfile(path):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                "pytorch3d_pointops_b200: libpointops_b200.so is missing and could not be built "
                f"({e}). Run `python -m pytorch3d_pointops_b200.build`. There is no CPU fallback."
            ) from e
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise RuntimeError(f"pytorch3d_pointops_b200: cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(status: int, what: str) -> None:
    """Non-zero status -> RuntimeError (the reference raises c10::Error -> RuntimeError)."""
    if status != 0:
        msg = load().pops_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what}: {msg} (status {status})")


def launch_count() -> int:
    return int(load().pops_launch_count())
