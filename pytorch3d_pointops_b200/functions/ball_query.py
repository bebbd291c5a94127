"""ball_query -- host side (reference: functions/ball_query.py:55-142)."""
from typing import Union

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _C
from .knn import _KNN
from .utils import masked_gather


class _ball_query(Function):
    """forward = _C.ball_query; backward reuses the KNN backward with norm 2, relying on its
    skip of idx == -1 (reference: ball_query.py:37-52)."""

    @staticmethod
    def forward(ctx, p1, p2, lengths1, lengths2, K, radius):
        idx, dists = _C.ball_query(p1, p2, lengths1, lengths2, K, radius)
        ctx.save_for_backward(p1, p2, lengths1, lengths2, idx)
        ctx.mark_non_differentiable(idx)
        return dists, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_dists, grad_idx):
        p1, p2, lengths1, lengths2, idx = ctx.saved_tensors
        if grad_dists.dtype != torch.float32:
            grad_dists = grad_dists.float()
        grad_p1, grad_p2 = _C.knn_points_backward(
            p1.float(), p2.float(), lengths1, lengths2, idx, 2, grad_dists.contiguous()
        )
        return grad_p1, grad_p2, None, None, None, None


def ball_query(
    p1: torch.Tensor,
    p2: torch.Tensor,
    lengths1: Union[torch.Tensor, None] = None,
    lengths2: Union[torch.Tensor, None] = None,
    K: int = 500,
    radius: float = 0.2,
    return_nn: bool = True,
):
    """For every point of p1 (N,P1,D): the first K points of p2 (N,P2,D), in index order, that
    lie strictly within `radius`.  Returns the `_KNN` tuple (dists (N,P1,K) zero padded,
    idx (N,P1,K) int64 padded with -1, knn (N,P1,K,D) zero padded when `return_nn`).
    Same arguments, defaults and errors as the reference."""
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
    dists, idx = _ball_query.apply(p1, p2, lengths1, lengths2, K, radius)
    points_nn = masked_gather(p2, idx) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=points_nn)
