"""Python face of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Native loops live in oracle/pointops_oracle.c (plain C restatement of the reference's
csrc/<op>/<op>_cpu.cpp, loaded with ctypes).  The torch post-processing the reference does
in functions/*.py is restated here on CPU tensors, each function citing the reference lines
it follows.  Everything takes and returns CPU torch tensors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpointops_oracle.so")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pointops_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f(t: torch.Tensor):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), _f32p)


def _l(t: torch.Tensor):
    assert t.dtype == torch.int64 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), _i64p)


def _prep(p1, p2, lengths1, lengths2):
    p1 = p1.detach().cpu().float().contiguous()
    p2 = p2.detach().cpu().float().contiguous()
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.int64)
    if lengths2 is None:
        lengths2 = torch.full((N,), P2, dtype=torch.int64)
    return p1, p2, lengths1.cpu().long().contiguous(), lengths2.cpu().long().contiguous()


# --------------------------------------------------------------------------------------
# _C-level restatements (reference: csrc/ext.cpp:15-27)
# --------------------------------------------------------------------------------------
def knn_points_idx(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, q0=0, q1=-1, threads=1):
    """_C.knn_points_idx on CPU (knn_cpu.cpp:13-69).  Returns (idx, dists) -- that order."""
    p1, p2, l1, l2 = _prep(p1, p2, lengths1, lengths2)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    idx = torch.empty((N, P1, K), dtype=torch.int64)
    dists = torch.empty((N, P1, K), dtype=torch.float32)
    lib().oracle_knn_idx(_f(p1), _f(p2), _l(l1), _l(l2), _i64(N), _i64(P1), _i64(P2), _i64(D),
                         _i64(K), int(norm), _i64(q0), _i64(q1), _l(idx), _f(dists), int(threads))
    return idx, dists


def knn_points_backward(p1, p2, lengths1, lengths2, idx, norm, grad_dists):
    """_C.knn_points_backward on CPU (knn_cpu.cpp:75-128).  Returns (grad_p1, grad_p2)."""
    p1, p2, l1, l2 = _prep(p1, p2, lengths1, lengths2)
    idx = idx.cpu().long().contiguous()
    g = grad_dists.detach().cpu().float().contiguous()
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    K = idx.shape[2]
    g1 = torch.empty_like(p1)
    g2 = torch.empty_like(p2)
    lib().oracle_knn_backward(_f(p1), _f(p2), _l(l1), _l(l2), _l(idx), _f(g), _i64(N), _i64(P1),
                              _i64(P2), _i64(D), _i64(K), int(norm), _f(g1), _f(g2))
    return g1, g2


def sample_pdf_(bins, weights, outputs, eps):
    """_C.sample_pdf on CPU (sample_pdf_cpu.cpp:24-142): in place on `outputs` (B, n_samples)."""
    bins = bins.detach().cpu().float().contiguous()
    weights = weights.detach().cpu().float().contiguous()
    assert outputs.dtype == torch.float32 and outputs.is_contiguous() and outputs.device.type == "cpu"
    B, n_bins = weights.shape
    st = lib().oracle_sample_pdf(_f(bins), _f(weights), _f(outputs), _i64(B), _i64(n_bins), _i64(outputs.shape[1]),
                                 ctypes.c_float(eps))
    assert st == 0
    return outputs


def sample_pdf(bins, weights, n_samples, det=False, eps=1e-5, u=None):
    """functions/sample_pdf.py:14-66.  `u` (optional) supplies the uniform numbers so that both
    sides of a comparison draw the same ones."""
    n_bins = weights.shape[-1]
    batch_shape = bins.shape[:-1]
    if det:
        out = torch.linspace(0.0, 1.0, n_samples, dtype=torch.float32).expand(batch_shape + (n_samples,)).contiguous()
    else:
        out = (u if u is not None else torch.rand(batch_shape + (n_samples,))).detach().cpu().float().clone().contiguous()
    sample_pdf_(bins.reshape(-1, n_bins + 1), weights.reshape(-1, n_bins), out.view(-1, n_samples), eps)
    return out


def ball_query_idx(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, q0=0, q1=-1,
                   threads=1):
    """_C.ball_query on CPU (ball_query_cpu.cpp:12-54).  Returns (idx, dists)."""
    p1, p2, l1, l2 = _prep(p1, p2, lengths1, lengths2)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    idx = torch.empty((N, P1, K), dtype=torch.int64)
    dists = torch.empty((N, P1, K), dtype=torch.float32)
    lib().oracle_ball_query(_f(p1), _f(p2), _l(l1), _l(l2), _i64(N), _i64(P1), _i64(P2), _i64(D),
                            _i64(K), ctypes.c_float(radius), _i64(q0), _i64(q1), _l(idx),
                            _f(dists), int(threads))
    return idx, dists


def sample_farthest_points_idx(points, lengths, K, start_idxs, n0=0, n1=-1, threads=1):
    """_C.sample_farthest_points on CPU (sample_farthest_points_cpu.cpp:14-103)."""
    points = points.detach().cpu().float().contiguous()
    N, P, D = points.shape
    lengths = lengths.cpu().long().contiguous()
    K = K.cpu().long().contiguous()
    start_idxs = start_idxs.cpu().long().contiguous()
    max_K = int(K.max()) if N > 0 else 0
    out = torch.empty((N, max_K), dtype=torch.int64)
    lib().oracle_fps(_f(points), _l(lengths), _l(K), _l(start_idxs), _i64(N), _i64(P), _i64(D),
                     _i64(max_K), _i64(n0), _i64(n1), _l(out), int(threads))
    return out


def packed_to_padded_C(inputs_packed, first_idxs, max_size):
    """_C.packed_to_padded on CPU (packed_to_padded_tensor_cpu.cpp:11-40)."""
    x = inputs_packed.detach().cpu().float().contiguous()
    f = first_idxs.cpu().long().contiguous()
    Fn, D = x.shape
    B = f.shape[0]
    out = torch.empty((B, max_size, D), dtype=torch.float32)
    lib().oracle_packed_to_padded(_f(x), _l(f), _i64(Fn), _i64(B), _i64(max_size), _i64(D), _f(out))
    return out


def padded_to_packed_C(inputs_padded, first_idxs, num_inputs):
    """_C.padded_to_packed on CPU (packed_to_padded_tensor_cpu.cpp:42-70)."""
    x = inputs_padded.detach().cpu().float().contiguous()
    f = first_idxs.cpu().long().contiguous()
    B, M, D = x.shape
    out = torch.empty((num_inputs, D), dtype=torch.float32)
    lib().oracle_padded_to_packed(_f(x), _l(f), _i64(num_inputs), _i64(B), _i64(M), _i64(D), _f(out))
    return out


# --------------------------------------------------------------------------------------
# functions/*.py restatements (CPU torch)
# --------------------------------------------------------------------------------------
def knn_gather(x, idx, lengths=None):
    """functions/knn.py:200-250: x_out[n,l,k] = x[n, idx[n,l,k]], zero where k >= lengths[n]."""
    N, M, U = x.shape
    _, L, K = idx.shape
    if lengths is None:
        lengths = torch.full((N,), M, dtype=torch.int64)
    out = x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
    if lengths.min() < K:
        mask = lengths[:, None] <= torch.arange(K)[None]
        out = out.masked_fill(mask[:, None, :, None].expand(-1, L, -1, U), 0.0)
    return out


def masked_gather(points, idx):
    """functions/utils.py:20-65: gather where idx == -1 yields a zero row."""
    N, P, D = points.shape
    mask = idx.eq(-1)
    safe = idx.clamp(min=0)
    if idx.ndim == 3:
        K = idx.shape[2]
        out = points[:, :, None, :].expand(-1, -1, K, -1).gather(
            1, safe[..., None].expand(-1, -1, -1, D))
    else:
        out = points.gather(1, safe[..., None].expand(-1, -1, D))
    return out.masked_fill(mask[..., None].expand_as(out), 0.0)


class _OracleKnn(torch.autograd.Function):
    """functions/knn.py:21-111 on the oracle's native loops (canonical, already sorted)."""

    @staticmethod
    def forward(ctx, p1, p2, lengths1, lengths2, K, norm):
        idx, dists = knn_points_idx(p1, p2, lengths1, lengths2, norm, K)
        ctx.save_for_backward(p1, p2, lengths1, lengths2, idx)
        ctx.mark_non_differentiable(idx)
        ctx.norm = norm
        return dists, idx

    @staticmethod
    def backward(ctx, grad_dists, grad_idx):
        p1, p2, lengths1, lengths2, idx = ctx.saved_tensors
        g1, g2 = knn_points_backward(p1, p2, lengths1, lengths2, idx, ctx.norm, grad_dists)
        return g1, g2, None, None, None, None


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, return_nn=False):
    """functions/knn.py:114-197 with the canonical (dist, idx)-lexicographic order.

    The reference's post-sort (knn.py:77-89) is the identity on the CPU kernel's already
    sorted output except for the order inside exact-tie groups when K > 16;
    the canonical raw order is the contract.
    """
    N, P1, _ = p1.shape
    P2 = p2.shape[1]
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.int64)
    if lengths2 is None:
        lengths2 = torch.full((N,), P2, dtype=torch.int64)
    dists, idx = _OracleKnn.apply(p1, p2, lengths1, lengths2, K, norm)
    nn = knn_gather(p2, idx, lengths2) if return_nn else None
    return dists, idx, nn


def ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    """functions/ball_query.py:55-142 (forward only)."""
    idx, dists = ball_query_idx(p1, p2, lengths1, lengths2, K, radius)
    nn = masked_gather(p2.detach().cpu().float(), idx) if return_nn else None
    return dists, idx, nn


def sample_farthest_points(points, lengths=None, K=50, start_idxs=None):
    """functions/sample_farthest_points.py:18-96 with explicit start indices."""
    N, P, D = points.shape
    if lengths is None:
        lengths = torch.full((N,), P, dtype=torch.int64)
    if isinstance(K, int):
        K = torch.full((N,), K, dtype=torch.int64)
    elif isinstance(K, list):
        K = torch.tensor(K, dtype=torch.int64)
    if start_idxs is None:
        start_idxs = torch.zeros_like(lengths)
    idx = sample_farthest_points_idx(points, lengths, K, start_idxs)
    return masked_gather(points.detach().cpu().float(), idx), idx


def _chamfer_single_direction(x, y, x_lengths, y_lengths, x_features, y_features, weights,
                              point_reduction, norm, abs_cosine, feature_names):
    """functions/chamfer.py:85-189."""
    return_features = (x_features is not None and y_features is not None
                       and feature_names is not None and len(feature_names) > 0)
    N, P1, D = x.shape
    x_mask = torch.arange(P1)[None] >= x_lengths[:, None]
    if weights is not None and weights.sum() == 0.0:
        w = weights.view(N, 1)
        z = (x.sum((1, 2)) * w) * 0.0
        return (z, z)
    dists, idx, _ = knn_points(x, y, x_lengths, y_lengths, norm=norm, K=1)
    cham_x = dists[..., 0].masked_fill(x_mask, 0.0)
    if weights is not None:
        cham_x = cham_x * weights.view(N, 1)
    cham_feat = None
    if return_features:
        cham_feat = {}
        for name in feature_names:
            near = knn_gather(y_features[name], idx, y_lengths)[..., 0, :]
            cos = F.cosine_similarity(x_features[name], near, dim=2, eps=1e-6)
            cos = cos.abs() if abs_cosine else cos
            fd = (1 - cos).masked_fill(x_mask, 0.0)
            if weights is not None:
                fd = fd * weights.view(N, 1)
            cham_feat[name] = fd
    if point_reduction == "max":
        cham_x = cham_x.max(1).values
    elif point_reduction is not None:
        cham_x = cham_x.sum(1)
        if return_features:
            cham_feat = {k: v.sum(1) for k, v in cham_feat.items()}
        if point_reduction == "mean":
            clamped = x_lengths.clamp(min=1)
            cham_x = cham_x / clamped
            if return_features:
                cham_feat = {k: v / clamped for k, v in cham_feat.items()}
    return cham_x, cham_feat


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_features=None, y_features=None,
                     weights=None, batch_reduction="mean", point_reduction="mean", norm=2,
                     single_directional=False, abs_cosine=True, feature_names=None):
    """functions/chamfer.py:217-365 on padded CPU tensors (differentiable through autograd)."""
    N = x.shape[0]
    if x_lengths is None:
        x_lengths = torch.full((N,), x.shape[1], dtype=torch.int64)
    if y_lengths is None:
        y_lengths = torch.full((N,), y.shape[1], dtype=torch.int64)
    cx, fx = _chamfer_single_direction(x, y, x_lengths, y_lengths, x_features, y_features,
                                       weights, point_reduction, norm, abs_cosine, feature_names)
    if single_directional:
        loss, lf = cx, fx
    else:
        cy, fy = _chamfer_single_direction(y, x, y_lengths, x_lengths, y_features, x_features,
                                           weights, point_reduction, norm, abs_cosine,
                                           feature_names)
        if point_reduction == "max":
            loss, lf = torch.maximum(cx, cy), None
        elif point_reduction is not None:
            loss = cx + cy
            lf = None if fx is None else {k: fx[k] + fy[k] for k in fx}
        else:
            loss = (cx, cy)
            lf = None if fx is None else {k: (fx[k], fy[k]) for k in fx}
    if batch_reduction is None:
        return loss, lf
    # chamfer.py:192-214
    loss = loss.sum()
    if lf is not None:
        lf = {k: v.sum() for k, v in lf.items()}
    if batch_reduction == "mean":
        if weights is None:
            div = max(N, 1)
        elif weights.sum() == 0.0:
            div = 1
        else:
            div = weights.sum()
        loss = loss / div
        if lf is not None:
            lf = {k: v / div for k, v in lf.items()}
    return loss, lf


def as_numpy(t: Optional[torch.Tensor]):
    return None if t is None else t.detach().cpu().numpy()


__all__ = [n for n in dir() if not n.startswith("_")] + ["np"]
