"""knn_points / knn_gather -- host side of the KNN hot path.

Mirrors the public API of the reference's functions/knn.py (knn_points :114-197, knn_gather
:200-250, the `_KNN` namedtuple :18) on top of the sm_100a kernels:

* the kernel already returns the canonical result (K lexicographically smallest (dist, idx),
  ascending), so the reference's `lengths2.min()` host sync + sort + gather post-pass
  (knn.py:77-89) does not exist here; `return_sorted` is accepted and has no cost;
* `knn_gather` is one fused gather kernel (idx read once, no expanded temporaries) with a
  scatter-add backward, instead of expand + gather + masked write (knn.py:233-248).
"""
from collections import namedtuple
from typing import Union

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _C

_KNN = namedtuple("KNN", "dists idx knn")


class _knn_points(Function):
    """autograd wrapper: forward = _C.knn_points_idx, backward = _C.knn_points_backward."""

    @staticmethod
    def forward(ctx, p1, p2, lengths1, lengths2, K, version, norm: int = 2,
                return_sorted: bool = True):
        if not ((norm == 1) or (norm == 2)):
            raise ValueError("Support for 1 or 2 norm.")
        idx, dists = _C.knn_points_idx(p1, p2, lengths1, lengths2, norm, K, version)
        ctx.save_for_backward(p1, p2, lengths1, lengths2, idx)
        ctx.mark_non_differentiable(idx)
        ctx.norm = norm
        return dists, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_dists, grad_idx):
        p1, p2, lengths1, lengths2, idx = ctx.saved_tensors
        if grad_dists.dtype != torch.float32:
            grad_dists = grad_dists.float()
        grad_p1, grad_p2 = _C.knn_points_backward(
            p1.float(), p2.float(), lengths1, lengths2, idx, ctx.norm, grad_dists.contiguous()
        )
        return grad_p1, grad_p2, None, None, None, None, None, None


class _gather_rows(Function):
    """x (N,M,U), idx (N,L,K) -> (N,L,K,U); mode = _C.GATHER_KNN | _C.GATHER_MASKED."""

    @staticmethod
    def forward(ctx, x, idx, lengths, mode, oob_flag):
        out = _C.gather(x, idx, lengths, mode, oob_flag)
        ctx.save_for_backward(idx, lengths if lengths is not None else idx.new_empty(0))
        ctx.has_lengths = lengths is not None
        ctx.M = x.shape[1]
        ctx.mode = mode
        ctx.mark_non_differentiable(idx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        idx, lengths = ctx.saved_tensors
        grad_x = _C.gather_backward(grad_out.float().contiguous(), idx,
                                    lengths if ctx.has_lengths else None, ctx.M, ctx.mode)
        return grad_x, None, None, None, None


def _gather_any(x, idx, lengths, mode, flag):
    """Row gather for any dtype: float32 goes straight to the kernel (differentiable); other
    dtypes are pure copies, so they are reinterpreted as float32 words when that is possible
    without autograd, else computed in float32 and cast back."""
    x = x.contiguous()
    idx = idx.contiguous()
    if x.dtype == torch.float32:
        return _gather_rows.apply(x, idx, lengths, mode, flag)
    needs_grad = x.requires_grad and torch.is_grad_enabled()
    nbytes = x.element_size() * x.shape[-1]
    if not needs_grad and nbytes % 4 == 0 and x.dtype != torch.bool:
        out = _gather_rows.apply(x.view(torch.float32), idx, lengths, mode, flag)
        return out.view(x.dtype)
    return _gather_rows.apply(x.float(), idx, lengths, mode, flag).to(x.dtype)


def knn_points(
    p1: torch.Tensor,
    p2: torch.Tensor,
    lengths1: Union[torch.Tensor, None] = None,
    lengths2: Union[torch.Tensor, None] = None,
    norm: int = 2,
    K: int = 1,
    version: int = -1,
    return_nn: bool = False,
    return_sorted: bool = True,
) -> _KNN:
    """K nearest neighbours of every point of p1 (N,P1,D) in p2 (N,P2,D).

    Same arguments and return value as the reference (functions/knn.py:114-197): `dists`
    (N,P1,K) squared L2 (or L1) distances, `idx` (N,P1,K) int64, both zero-padded where
    `k >= lengths2[n]` or `i >= lengths1[n]`; `knn` (N,P1,K,D) when `return_nn`.  Results are
    always sorted ascending by (dist, idx); `version` is accepted and ignored.
    """
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    p1 = p1.contiguous()
    p2 = p2.contiguous()
    N, P1 = p1.shape[0], p1.shape[1]
    P2 = p2.shape[1]
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.int64, device=p1.device)
    if lengths2 is None:
        lengths2 = torch.full((N,), P2, dtype=torch.int64, device=p1.device)

    p1_dists, p1_idx = _knn_points.apply(p1, p2, lengths1, lengths2, K, version, norm, return_sorted)

    p2_nn = None
    if return_nn:
        # indices come from our own kernel: always in range, no bounds flag needed
        p2_nn = _gather_rows.apply(p2, p1_idx, lengths2, _C.GATHER_KNN, None)
    return _KNN(dists=p1_dists, idx=p1_idx, knn=p2_nn)


def knn_gather(x: torch.Tensor, idx: torch.Tensor, lengths: Union[torch.Tensor, None] = None):
    """x_out[n,l,k] = x[n, idx[n,l,k]] for x (N,M,U), idx (N,L,K); zero where k >= lengths[n].

    Reference: functions/knn.py:200-250.  An index outside [0, M) (e.g. the -1 padding of
    ball_query) is an error there (`RuntimeError: index -1 is out of bounds`); the same
    RuntimeError is raised here (one 4-byte device->host read) -- use `masked_gather` for
    -1-padded indices.
    """
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    flag = torch.zeros((1,), dtype=torch.int32, device=x.device)
    out = _gather_any(x, idx, lengths, _C.GATHER_KNN, flag)
    if int(flag.item()) != 0:
        raise RuntimeError(f"index out of bounds in knn_gather: idx must lie in [0, {M})")
    return out
