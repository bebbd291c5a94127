"""sample_farthest_points (+ the pure-torch naive variant) -- reference:
functions/sample_farthest_points.py:18-197."""
from random import randint
from typing import List, Optional, Tuple, Union

import torch

from .. import _C
from .utils import masked_gather


def _normalise_k(K, N, device):
    """int | list | tensor -> (int64 tensor (N,), max_K or None when only known on device)."""
    if isinstance(K, int):
        return torch.full((N,), K, dtype=torch.int64, device=device), K
    if isinstance(K, list):
        return torch.tensor(K, dtype=torch.int64, device=device), (max(K) if len(K) else 0)
    return K, None


def sample_farthest_points(
    points: torch.Tensor,
    lengths: Optional[torch.Tensor] = None,
    K: Union[int, List, torch.Tensor] = 50,
    random_start_point: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Iterative farthest point sampling of K points per cloud from points (N,P,D).

    Returns (selected_points (N,max K,D) zero padded, selected_indices (N,max K) int64 padded
    with -1).  Arguments, dtype coercions and errors follow the reference (:55-96); the
    selection itself is not differentiable, the returned points are (through the gather).
    When K is an int or a list no device->host sync is needed (the reference always syncs on
    max(K), sample_farthest_points.cu:132).
    """
    N, P, D = points.shape
    device = points.device
    if lengths is None:
        lengths = torch.full((N,), P, dtype=torch.int64, device=device)
    else:
        if lengths.shape != (N,):
            raise ValueError("points and lengths must have same batch dimension.")
        if lengths.max() > P:
            raise ValueError("A value in lengths was too large.")
    K, max_K = _normalise_k(K, N, device)
    if K.shape[0] != N:
        raise ValueError("K and points must have the same batch dimension")
    if points.dtype != torch.float32:
        points = points.to(torch.float32)
    if lengths.dtype != torch.int64:
        lengths = lengths.to(torch.int64)
    if K.dtype != torch.int64:
        K = K.to(torch.int64)

    start_idxs = torch.zeros_like(lengths)
    if random_start_point:
        # same CPU-generator draws as the reference (:87-89), one host copy of lengths
        highs = lengths.tolist()
        start_idxs = torch.tensor(
            [int(torch.randint(high=h, size=(1,)).item()) for h in highs],
            dtype=torch.int64, device=device)

    with torch.no_grad():
        idx = _C.sample_farthest_points(points.contiguous(), lengths, K, start_idxs, max_K)
    return masked_gather(points, idx), idx


def sample_farthest_points_naive(
    points: torch.Tensor,
    lengths: Optional[torch.Tensor] = None,
    K: Union[int, List, torch.Tensor] = 50,
    random_start_point: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pure-torch farthest point sampling, one cloud at a time (reference :99-197).  Kept as
    the API's second, device-agnostic implementation; not a fallback of the CUDA path."""
    N, P, D = points.shape
    device = points.device
    if lengths is None:
        lengths = torch.full((N,), P, dtype=torch.int64, device=device)
    else:
        if lengths.shape != (N,):
            raise ValueError("points and lengths must have same batch dimension.")
        if lengths.max() > P:
            raise ValueError("Invalid lengths.")
    K, _ = _normalise_k(K, N, device)
    if K.shape[0] != N:
        raise ValueError("K and points must have the same batch dimension")
    max_K = int(torch.max(K))
    rows = []
    for n in range(N):
        row = torch.full((max_K,), -1, dtype=torch.int64, device=device)
        ln = int(lengths[n])
        closest = points.new_full((ln,), float("inf"), dtype=torch.float32)
        cur = randint(0, ln - 1) if random_start_point else 0
        row[0] = cur
        for i in range(1, min(ln, int(K[n]))):
            delta = points[n, cur, :] - points[n, :ln, :]
            closest = torch.min((delta**2).sum(-1), closest)
            cur = torch.argmax(closest)
            row[i] = cur
        rows.append(row)
    all_idx = torch.stack(rows, dim=0)
    if points.is_cuda:
        return masked_gather(points, all_idx), all_idx
    safe = all_idx.clamp(min=0)
    pts = points.gather(1, safe[..., None].expand(-1, -1, D))
    return pts.masked_fill(all_idx.eq(-1)[..., None], 0.0), all_idx
