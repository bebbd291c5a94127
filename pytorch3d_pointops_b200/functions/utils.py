"""masked_gather, wmean, get_point_covariances (reference: functions/utils.py)."""
from typing import Optional, Tuple, Union

import torch

from .. import _C
from .knn import _gather_any, knn_points


def masked_gather(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Gather rows of `points` (N,P,D) at `idx` ((N,K) or (N,P',K) int64), where -1 marks
    padding and yields a row of zeros (reference: functions/utils.py:20-65).

    One fused kernel: the index is read once and no (N,P',K,D) int64 index expansion or
    masked rewrites are materialised.
    """
    if len(idx) != len(points):
        raise ValueError("points and idx must have the same batch dimension")
    if idx.ndim == 3:
        return _gather_any(points, idx, None, _C.GATHER_MASKED, None)
    if idx.ndim == 2:
        out = _gather_any(points, idx[:, None, :], None, _C.GATHER_MASKED, None)
        return out[:, 0]
    raise ValueError("idx format is not supported %s" % repr(idx.shape))


def wmean(
    x: torch.Tensor,
    weight: Optional[torch.Tensor] = None,
    dim: Union[int, Tuple[int]] = -2,
    keepdim: bool = True,
    eps: float = 1e-9,
) -> torch.Tensor:
    """(Weighted) mean of x (*, D) over `dim`: sum(x*w) / max(sum(w), eps)
    (reference: functions/utils.py:68-108)."""
    if weight is None:
        return x.mean(dim=dim, keepdim=keepdim)
    for xd, wd in zip(x.shape[-2::-1], weight.shape[::-1]):
        if xd != wd and xd != 1 and wd != 1:
            raise ValueError("wmean: weights are not compatible with the tensor")
    w = weight[..., None]
    return (x * w).sum(dim=dim, keepdim=keepdim) / w.sum(dim=dim, keepdim=keepdim).clamp(eps)


def get_point_covariances(
    points_padded: torch.Tensor,
    num_points_per_cloud: torch.Tensor,
    neighborhood_size: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-point covariance of the K nearest neighbours (reference: functions/utils.py:111-153).
    Returns (covariances (N,P,D,D), k_nearest_neighbors (N,P,K,D))."""
    D = points_padded.shape[-1]
    if 1 <= D <= 4 and not (torch.is_grad_enabled() and points_padded.requires_grad):
        # no gradient wanted: search, then ONE kernel gathers the neighbourhoods and reduces them to
        # covariances (no (N,P,K,D,D) outer-product temporary)
        res = knn_points(points_padded, points_padded, lengths1=num_points_per_cloud,
                         lengths2=num_points_per_cloud, K=neighborhood_size, return_nn=False)
        return _C.point_covariances(points_padded.float(), res.idx, num_points_per_cloud)
    nn = knn_points(
        points_padded,
        points_padded,
        lengths1=num_points_per_cloud,
        lengths2=num_points_per_cloud,
        K=neighborhood_size,
        return_nn=True,
    ).knn
    centered = nn - nn.mean(2, keepdim=True)
    cov = (centered.unsqueeze(4) * centered.unsqueeze(3)).mean(2)
    return cov, nn
