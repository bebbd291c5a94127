"""chamfer_distance with per-feature cosine losses (reference: functions/chamfer.py:217-365).

Same signature, validation, return structure and numerics contract as the reference.  Each
direction is one autograd Function of two fused kernels around the K=1 search of the sm_100a KNN
kernel (mask, weights, neighbour-feature gather, cosine terms and point reduction in one pass;
distance and cosine gradients with their index scatters in another).
"""
from typing import Union

import torch

from ..structures.point_structure import Pointclouds
from .. import _C


def _validate_chamfer_reduction_inputs(
    batch_reduction: Union[str, None], point_reduction: Union[str, None]
) -> None:
    """batch_reduction in {"mean","sum",None}; point_reduction in {"mean","sum","max",None};
    a per-point result (point_reduction None) cannot be batch-reduced (reference :17-35)."""
    if batch_reduction is not None and batch_reduction not in ["mean", "sum"]:
        raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
    if point_reduction is not None and point_reduction not in ["mean", "sum", "max"]:
        raise ValueError('point_reduction must be one of ["mean", "sum", "max"] or None')
    if point_reduction is None and batch_reduction is not None:
        raise ValueError("Batch reduction must be None if point_reduction is None")


def _handle_pointcloud_input(
    points: Union[torch.Tensor, Pointclouds],
    lengths: Union[torch.Tensor, None],
    features: Union[torch.Tensor, dict, None],
):
    """Pointclouds -> (padded points, num_points_per_cloud, padded feature dict); a tensor is
    validated and passed through with default full lengths (reference :38-82)."""
    if isinstance(points, Pointclouds):
        return points.points_padded(), points.num_points_per_cloud(), points.features_padded()
    if not torch.is_tensor(points):
        raise ValueError(
            "The input pointclouds should be either "
            + "Pointclouds objects or torch.Tensor of shape "
            + "(minibatch, num_points, 3)."
        )
    if points.ndim != 3:
        raise ValueError("Expected points to be of shape (N, P, D)")
    if lengths is not None:
        if lengths.ndim != 1 or lengths.shape[0] != points.shape[0]:
            raise ValueError("Expected lengths to be of shape (N,)")
        if _lengths_too_long(lengths, points.shape[1]):
            raise ValueError("A length value was too long")
    else:
        lengths = torch.full((points.shape[0],), points.shape[1], dtype=torch.int64,
                             device=points.device)
    if isinstance(features, dict):
        for name, tensor in features.items():
            if tensor is not None and tensor.ndim != 3:
                raise ValueError(f"Expected {name} to be of shape (N, P, C)")
    elif torch.is_tensor(features) and features.ndim != 3:
        raise ValueError("Expected features to be of shape (N, P, C)")
    return points, lengths, features


# lengths tensors already checked against a padded size: the same tensor object at the same version
# need not pay the device->host read again (the check itself is the reference's, :58-60)
_LENGTHS_OK = {}


def _lengths_too_long(lengths: torch.Tensor, P: int) -> bool:
    key = id(lengths)
    hit = _LENGTHS_OK.get(key)
    if hit is not None and hit[0]() is lengths and hit[1] == lengths._version and hit[2] <= P:
        return False
    too_long = bool(lengths.max() > P)
    if not too_long:
        import weakref

        if len(_LENGTHS_OK) > 256:
            _LENGTHS_OK.clear()
        _LENGTHS_OK[key] = (weakref.ref(lengths, lambda _r, k=key: _LENGTHS_OK.pop(k, None)), lengths._version, P)
    return too_long


class _ChamferDirection(torch.autograd.Function):
    """x -> y half of the chamfer loss as two fused kernels around the K=1 search.

    forward : _C.knn_points_idx (K=1) + _C.chamfer_forward (mask, weights, neighbour-feature
              gather, cosine terms, point reduction);
    backward: _C.chamfer_backward (distance gradient to x and scatter to y[idx], cosine chain rule
              to x features and scatter to y features[idx]).
    Outputs: cham (N,) [(N,P1) when point_reduction is None], then one tensor per feature.
    """

    @staticmethod
    def forward(ctx, x, y, x_lengths, y_lengths, weights, norm, point_reduction, abs_cosine, nfeat,
                *feats):
        xfs, yfs = list(feats[:nfeat]), list(feats[nfeat:])
        idx, dists = _C.knn_points_idx(x, y, x_lengths, y_lengths, norm, 1, -1)
        N, P1 = x.shape[0], x.shape[1]
        cham, fo, argmax = _C.chamfer_forward(dists.view(N, P1), idx.view(N, P1), x_lengths, y_lengths,
                                              weights, y.shape[1], xfs, yfs, point_reduction, abs_cosine)
        ctx.save_for_backward(x, y, x_lengths, y_lengths, idx, *xfs, *yfs)
        ctx.weights, ctx.argmax = weights, argmax
        ctx.cfg = (norm, point_reduction, abs_cosine, nfeat)
        return (cham,) + tuple(fo[f] for f in range(nfeat))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_cham, *g_feats):
        norm, point_reduction, abs_cosine, nfeat = ctx.cfg
        x, y, x_lengths, y_lengths, idx = ctx.saved_tensors[:5]
        xfs = list(ctx.saved_tensors[5:5 + nfeat])
        yfs = list(ctx.saved_tensors[5 + nfeat:])
        N, P1 = x.shape[0], x.shape[1]
        g_feat = None
        if nfeat:
            g_feat = torch.stack([g if g is not None else torch.zeros_like(g_cham) for g in g_feats], 0)
        gx, gy, gxf, gyf = _C.chamfer_backward(x, y, idx.view(N, P1), x_lengths, y_lengths, ctx.weights,
                                               norm, xfs, yfs, point_reduction, abs_cosine,
                                               g_cham.contiguous(), g_feat, ctx.argmax)
        return (gx, gy, None, None, None, None, None, None, None) + tuple(gxf) + tuple(gyf)


class _ChamferBoth(torch.autograd.Function):
    """Both directions of the loss for point_reduction "sum" / "mean" in one autograd node.

    Same kernels as two `_ChamferDirection` nodes; what it removes is the glue around them, which
    is what a step of this size spends its time on (the step is launch-bound: ~65 launches of a
    few microseconds): the two directions write into ONE (2, 1+F, N) tensor, one `sum` adds the
    directions (and the batch), and the backward runs the y -> x kernel in accumulate mode on top
    of the x -> y gradients -- no per-tensor zero fills, no `grad_a + grad_b`.
    Outputs: (1+F) tensors -- the loss and one per feature -- shaped (N,) or () after the batch
    reduction (`batch_mode` 0 none | 1 sum | 2 mean over N clouds)."""

    @staticmethod
    def forward(ctx, x, y, x_lengths, y_lengths, norm, point_reduction, abs_cosine, nfeat, batch_mode, *feats):
        xfs, yfs = list(feats[:nfeat]), list(feats[nfeat:])
        N, P1, P2 = x.shape[0], x.shape[1], y.shape[1]
        out = torch.empty((2, 1 + nfeat, N), dtype=torch.float32, device=x.device)
        idx1, d1, idx2, d2 = _C.knn_points_idx_pair(x, y, x_lengths, y_lengths, norm, 1)  # one pre-pass for both
        _C.chamfer_forward(d1.view(N, P1), idx1.view(N, P1), x_lengths, y_lengths, None, P2, xfs, yfs,
                           point_reduction, abs_cosine, out=out[0])
        _C.chamfer_forward(d2.view(N, P2), idx2.view(N, P2), y_lengths, x_lengths, None, P1, yfs, xfs,
                           point_reduction, abs_cosine, out=out[1])
        if batch_mode == 0:
            total = out.sum(0)  # (1+F, N)
        else:
            total = out.sum((0, 2))  # (1+F,)
            if batch_mode == 2:
                total = total / max(N, 1)
        ctx.save_for_backward(x, y, x_lengths, y_lengths, idx1, idx2, *xfs, *yfs)
        ctx.cfg = (norm, point_reduction, abs_cosine, nfeat, batch_mode)
        return tuple(total[i] for i in range(1 + nfeat))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        norm, point_reduction, abs_cosine, nfeat, batch_mode = ctx.cfg
        x, y, x_lengths, y_lengths, idx1, idx2 = ctx.saved_tensors[:6]
        xfs = list(ctx.saved_tensors[6:6 + nfeat])
        yfs = list(ctx.saved_tensors[6 + nfeat:])
        N, P1, P2 = x.shape[0], x.shape[1], y.shape[1]
        like = next(g for g in grads if g is not None)
        g = torch.stack([gi if gi is not None else torch.zeros_like(like) for gi in grads], 0).float().contiguous()
        # batch-reduced loss: the (1+F) upstream scalars serve every cloud (the kernel broadcasts and
        # scales them), no (1+F, N) expansion on the way
        bcast = batch_mode != 0
        g_scale = 1.0 / max(N, 1) if batch_mode == 2 else 1.0
        g_cham, g_feat = g[0:1] if bcast else g[0], (g[1:] if nfeat else None)
        # one zero fill for every gradient, carved into the per-tensor views
        srcs = [x, y] + xfs + yfs
        flat = torch.zeros(sum(t.numel() for t in srcs), dtype=torch.float32, device=x.device)
        views, off = [], 0
        for t in srcs:
            views.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        gx, gy, gxf, gyf = views[0], views[1], views[2:2 + nfeat], views[2 + nfeat:]
        _C.chamfer_backward(x, y, idx1.view(N, P1), x_lengths, y_lengths, None, norm, xfs, yfs, point_reduction,
                            abs_cosine, g_cham, g_feat, None, into=(gx, gy, gxf, gyf), g_broadcast=bcast, g_scale=g_scale)
        _C.chamfer_backward(y, x, idx2.view(N, P2), y_lengths, x_lengths, None, norm, yfs, xfs, point_reduction,
                            abs_cosine, g_cham, g_feat, None, into=(gy, gx, gyf, gxf), g_broadcast=bcast, g_scale=g_scale)
        return (gx, gy) + (None,) * 7 + tuple(gxf) + tuple(gyf)


def _chamfer_distance_single_direction(
    x, y, x_lengths, y_lengths, x_features, y_features, weights,
    point_reduction: Union[str, None], norm: int, abs_cosine: bool,
    feature_names: Union[list, None] = None,
):
    """x -> y half of the loss (reference :85-189): NN distance of every x point, optional
    per-feature 1 - |cos| to the neighbour's feature, masks for ragged clouds, batch weights,
    and the reduction over points.  Validation as in the reference; the computation itself is
    `_ChamferDirection`."""
    if feature_names and x_features is not None and y_features is not None:
        for name in feature_names:
            if name not in x_features:
                raise ValueError(f"Feature '{name}' is missing in x_features.")
            if name not in y_features:
                raise ValueError(f"Feature '{name}' is missing in y_features.")
    with_features = (
        x_features is not None and y_features is not None
        and feature_names is not None and len(feature_names) > 0
    )
    N, P1, D = x.shape
    if y.shape[0] != N or y.shape[2] != D:
        raise ValueError("y does not have the correct shape.")
    if weights is not None:
        if weights.size(0) != N:
            raise ValueError("weights must be of shape (N,).")
        if not (weights >= 0).all():
            raise ValueError("weights cannot be negative.")
        if weights.sum() == 0.0:
            w = weights.view(N, 1)
            return ((x.sum((1, 2)) * w) * 0.0, (x.sum((1, 2)) * w) * 0.0)
    if point_reduction == "max":
        assert not with_features
    names = list(feature_names) if with_features else []
    if len(names) > 8:
        raise ValueError("at most 8 feature names are supported per call")
    xfs = [x_features[n].contiguous() for n in names]
    yfs = [y_features[n].contiguous() for n in names]
    # learned per-cloud weights: the reference's `cham_x *= weights.view(N, 1)` (:150-151, :175-176) is
    # differentiable in `weights`; the kernel takes them as constants, so in that case the per-cloud
    # results are computed unweighted and scaled by torch (w >= 0 commutes with sum / mean / max)
    w_grad = weights is not None and weights.requires_grad and torch.is_grad_enabled()
    out = _ChamferDirection.apply(
        x.contiguous(), y.contiguous(), x_lengths, y_lengths,
        None if (weights is None or w_grad) else weights.detach().float().contiguous(),
        norm, point_reduction, bool(abs_cosine), len(names), *xfs, *yfs)
    if w_grad:
        w = weights.view(N, 1) if point_reduction is None else weights.view(N)
        out = tuple(o * w for o in out)
    cham_feat = {n: out[1 + i] for i, n in enumerate(names)} if with_features else None
    return out[0], cham_feat


def _apply_batch_reduction(cham_x, cham_features_x, weights, batch_reduction: Union[str, None]):
    """Sum over the batch, divided by N (or sum of weights) for "mean" (reference :192-214)."""
    if batch_reduction is None:
        return (cham_x, cham_features_x)
    N = cham_x.shape[0]
    cham_x = cham_x.sum()
    if cham_features_x is not None:
        cham_features_x = {k: v.sum() for k, v in cham_features_x.items()}
    if batch_reduction == "mean":
        if weights is None:
            div = max(N, 1)
        elif weights.sum() == 0.0:
            div = 1
        else:
            div = weights.sum()
        cham_x = cham_x / div
        if cham_features_x is not None:
            cham_features_x = {k: v / div for k, v in cham_features_x.items()}
    return (cham_x, cham_features_x)


def _chamfer_both(x, y, x_lengths, y_lengths, x_features, y_features, weights, batch_reduction,
                  point_reduction, norm, single_directional, abs_cosine, feature_names):
    """The common two-sided, unweighted, point-reduced call as ONE autograd node (`_ChamferBoth`);
    None when the call is of another kind (the per-direction path then serves it).  Same
    validation and the same error messages as the per-direction path."""
    if single_directional or weights is not None or point_reduction not in ("mean", "sum"):
        return None
    with_features = (
        x_features is not None and y_features is not None
        and feature_names is not None and len(feature_names) > 0
    )
    if feature_names and x_features is not None and y_features is not None:
        for name in feature_names:
            if name not in x_features:
                raise ValueError(f"Feature '{name}' is missing in x_features.")
            if name not in y_features:
                raise ValueError(f"Feature '{name}' is missing in y_features.")
    N, P1, D = x.shape
    if y.shape[0] != N or y.shape[2] != D:
        raise ValueError("y does not have the correct shape.")
    names = list(feature_names) if with_features else []
    if len(names) > 8:
        raise ValueError("at most 8 feature names are supported per call")
    if N == 0:
        return None
    xfs = [x_features[n].contiguous() for n in names]
    yfs = [y_features[n].contiguous() for n in names]
    batch_mode = {None: 0, "sum": 1, "mean": 2}[batch_reduction]
    out = _ChamferBoth.apply(x.contiguous(), y.contiguous(), x_lengths, y_lengths, norm, point_reduction,
                             bool(abs_cosine), len(names), batch_mode, *xfs, *yfs)
    return out[0], ({n: out[1 + i] for i, n in enumerate(names)} if with_features else None)


def chamfer_distance(
    x,
    y,
    x_lengths=None,
    y_lengths=None,
    x_features=None,
    y_features=None,
    weights=None,
    batch_reduction: Union[str, None] = "mean",
    point_reduction: Union[str, None] = "mean",
    norm: int = 2,
    single_directional: bool = False,
    abs_cosine: bool = True,
    feature_names: Union[list, None] = None,
):
    """Chamfer distance between x (N,P1,D) and y (N,P2,D) (tensors or Pointclouds).

    Returns `(loss, loss_features)` exactly as the reference does (:272-286): reduced tensors
    for point_reduction in {"mean","sum","max"}, a (x->y, y->x) tuple of (N,P) tensors for
    point_reduction None (a bare tensor when single_directional); `loss_features` maps each
    name in `feature_names` to the cosine loss of that feature (None when no features).
    """
    _validate_chamfer_reduction_inputs(batch_reduction, point_reduction)
    if not ((norm == 1) or (norm == 2)):
        raise ValueError("Support for 1 or 2 norm.")
    if point_reduction == "max" and (feature_names is not None and len(feature_names) > 0):
        raise ValueError('Features must be None if point_reduction is "max"')

    x, x_lengths, x_features = _handle_pointcloud_input(x, x_lengths, x_features)
    y, y_lengths, y_features = _handle_pointcloud_input(y, y_lengths, y_features)

    fused = _chamfer_both(x, y, x_lengths, y_lengths, x_features, y_features, weights, batch_reduction,
                          point_reduction, norm, single_directional, abs_cosine, feature_names)
    if fused is not None:
        return fused

    cham_x, feat_x = _chamfer_distance_single_direction(
        x, y, x_lengths, y_lengths, x_features, y_features, weights, point_reduction, norm,
        abs_cosine, feature_names)
    if single_directional:
        loss, loss_features = cham_x, feat_x
    else:
        cham_y, feat_y = _chamfer_distance_single_direction(
            y, x, y_lengths, x_lengths, y_features, x_features, weights, point_reduction, norm,
            abs_cosine, feature_names)
        if point_reduction == "max":
            loss, loss_features = torch.maximum(cham_x, cham_y), None
        elif point_reduction is not None:
            loss = cham_x + cham_y
            loss_features = None
            if feat_x is not None:
                loss_features = {k: (feat_x[k] + feat_y[k]) if k in feat_y else feat_x[k] for k in feat_x}
        else:
            loss = (cham_x, cham_y)
            loss_features = None
            if feat_x is not None:
                loss_features = {k: (feat_x[k], feat_y.get(k)) for k in feat_x}
    return _apply_batch_reduction(loss, loss_features, weights, batch_reduction)
