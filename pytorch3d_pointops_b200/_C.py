"""`pytorch3d_pointops._C` replacement: the reference's pybind surface (csrc/ext.cpp:15-27) on top
of the C ABI of libpointops_b200.so.

Same names, same positional signatures, same return order and dtypes as the reference module
(`knn_points_idx` returns `(idx, dists)`; idx int64; KNN pads with 0, ball query / FPS with -1).
Inputs must be CUDA tensors: there is no CPU path (the reference's CPU loops survive only as the
test oracle).  Outputs are fresh tensors on the inputs' device; kernels are enqueued on the
current stream of that device and the call returns without synchronising (knn.cu:330-331).

Extra, additive entry points used by functions/*.py: `gather`, `gather_backward`, the fused
chamfer pieces.
"""
from __future__ import annotations

import torch

from . import _lib

GATHER_KNN = 0
GATHER_MASKED = 1


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        # reference: CHECK_CUDA -> "<name> must be a CUDA tensor." (pytorch3d_cutils.h:12)
        raise RuntimeError(f"{name} must be a CUDA tensor. (pytorch3d_pointops_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} for {name}")
    return t.contiguous()


def _cuda_i64(t: torch.Tensor, name: str, like: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor. (pytorch3d_pointops_b200 has no CPU path)")
    if t.device != like.device:
        raise RuntimeError(f"{name} must be on the same device as the points ({like.device})")
    if t.dtype != torch.int64:
        raise RuntimeError(f"expected scalar type Long but found {t.dtype} for {name}")
    return t.contiguous()


def _same_device(a: torch.Tensor, b: torch.Tensor, what: str) -> None:
    if a.device != b.device:
        raise RuntimeError(f"{what} must be on the same GPU")  # checkAllSameGPU, knn.cu:326


def _stream(t: torch.Tensor) -> int:
    # the raw handle of the current stream (what torch.cuda.current_stream(dev).cuda_stream returns,
    # without building two Python objects per call: these wrappers sit on launch-bound paths)
    return torch._C._cuda_getCurrentRawStream(t.device.index)


class _on_device:
    """`with torch.cuda.device(dev)` that costs nothing when `dev` is already current."""

    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        cur = torch._C._cuda_getDevice()
        if cur != self.idx:
            self.prev = cur
            torch._C._cuda_setDevice(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch._C._cuda_setDevice(self.prev)
        return False


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)


def _ptr(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# the reference's 7 in-scope names
# ------------------------------------------------------------------------------------------------
def knn_points_idx(p1, p2, lengths1, lengths2, norm, K, version):
    """knn.h:59-66.  Returns (idx (N,P1,K) int64, dists (N,P1,K) float32), canonical order."""
    lib = _lib.load()
    p1 = _cuda_f32(p1, "p1")
    p2 = _cuda_f32(p2, "p2")
    _same_device(p1, p2, "p1 and p2")
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[0] != p2.shape[0] or p1.shape[2] != p2.shape[2]:
        raise RuntimeError("p1 and p2 must be (N, P, D) with matching N and D")
    if norm not in (1, 2):
        raise RuntimeError("Norm must be 1 or 2.")  # knn.cu:339
    lengths1 = _cuda_i64(lengths1, "lengths1", p1)
    lengths2 = _cuda_i64(lengths2, "lengths2", p1)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    K = int(K)
    idx = torch.empty((N, P1, K), dtype=torch.int64, device=p1.device)
    dists = torch.empty((N, P1, K), dtype=torch.float32, device=p1.device)
    if idx.numel() == 0:
        return idx, dists
    with _on_device(p1.device):
        nbytes = lib.pops_knn_workspace_bytes(N, P1, P2, D, K, int(norm))
        ws = _ws(nbytes, p1.device)
        st = lib.pops_knn_points_idx(p1.data_ptr(), p2.data_ptr(), lengths1.data_ptr(),
                                     lengths2.data_ptr(), N, P1, P2, D, K, int(norm), int(version),
                                     idx.data_ptr(), dists.data_ptr(), ws.data_ptr(), ws.numel(),
                                     _stream(p1))
    _lib.check(st, "knn_points_idx")
    return idx, dists


def knn_points_idx_pair(p1, p2, lengths1, lengths2, norm, K):
    """pops_knn_points_idx_pair (additive): knn_points_idx(p1, p2, ...) and knn_points_idx(p2, p1, ...)
    in one call, sharing one spatial pre-pass.  Returns (idx12, dists12, idx21, dists21)."""
    lib = _lib.load()
    p1 = _cuda_f32(p1, "p1")
    p2 = _cuda_f32(p2, "p2")
    _same_device(p1, p2, "p1 and p2")
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[0] != p2.shape[0] or p1.shape[2] != p2.shape[2]:
        raise RuntimeError("p1 and p2 must be (N, P, D) with matching N and D")
    if norm not in (1, 2):
        raise RuntimeError("Norm must be 1 or 2.")
    lengths1 = _cuda_i64(lengths1, "lengths1", p1)
    lengths2 = _cuda_i64(lengths2, "lengths2", p1)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    K = int(K)
    dev = p1.device
    idx12 = torch.empty((N, P1, K), dtype=torch.int64, device=dev)
    d12 = torch.empty((N, P1, K), dtype=torch.float32, device=dev)
    idx21 = torch.empty((N, P2, K), dtype=torch.int64, device=dev)
    d21 = torch.empty((N, P2, K), dtype=torch.float32, device=dev)
    if N == 0 or K == 0:
        return idx12, d12, idx21, d21
    with _on_device(dev):
        ws = _ws(lib.pops_knn_pair_workspace_bytes(N, P1, P2, D, K, int(norm)), dev)
        st = lib.pops_knn_points_idx_pair(p1.data_ptr(), p2.data_ptr(), lengths1.data_ptr(), lengths2.data_ptr(),
                                          N, P1, P2, D, K, int(norm), idx12.data_ptr(), d12.data_ptr(),
                                          idx21.data_ptr(), d21.data_ptr(), ws.data_ptr(), ws.numel(), _stream(p1))
    _lib.check(st, "knn_points_idx_pair")
    return idx12, d12, idx21, d21


class KnnSliced:
    """pops_knn_points_prepare + pops_knn_points_idx_range (additive, pointops_b200.h): one pre-pass
    for the whole batch, then the rows of a range of clouds per call.  Owns idx / dists / workspace."""

    def __init__(self, p1, p2, lengths1, lengths2, norm, K):
        self.lib = _lib.load()
        self.p1 = _cuda_f32(p1, "p1")
        self.p2 = _cuda_f32(p2, "p2")
        _same_device(self.p1, self.p2, "p1 and p2")
        if norm not in (1, 2):
            raise RuntimeError("Norm must be 1 or 2.")
        self.l1 = _cuda_i64(lengths1, "lengths1", self.p1)
        self.l2 = _cuda_i64(lengths2, "lengths2", self.p1)
        self.N, self.P1, self.D = self.p1.shape
        self.P2 = self.p2.shape[1]
        self.K, self.norm = int(K), int(norm)
        dev = self.p1.device
        self.idx = torch.empty((self.N, self.P1, self.K), dtype=torch.int64, device=dev)
        self.dists = torch.empty((self.N, self.P1, self.K), dtype=torch.float32, device=dev)
        with _on_device(dev):
            self.ws = _ws(self.lib.pops_knn_workspace_bytes(self.N, self.P1, self.P2, self.D, self.K, self.norm), dev)

    def _args(self):
        return (self.p1.data_ptr(), self.p2.data_ptr(), self.l1.data_ptr(), self.l2.data_ptr(), self.N, self.P1,
                self.P2, self.D, self.K, self.norm)

    def prepare(self):
        with _on_device(self.p1.device):
            st = self.lib.pops_knn_points_prepare(*self._args(), self.ws.data_ptr(), self.ws.numel(), _stream(self.p1))
        _lib.check(st, "knn_points_prepare")

    def search(self, n0, n1):
        if self.idx.numel() == 0:
            return
        with _on_device(self.p1.device):
            st = self.lib.pops_knn_points_idx_range(*self._args(), -1, int(n0), int(n1), self.idx.data_ptr(),
                                                    self.dists.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                                    _stream(self.p1))
        _lib.check(st, "knn_points_idx_range")


def knn_check_version(version, D, K):
    """knn.h:161 / knn.cu:292-303."""
    return bool(_lib.load().pops_knn_check_version(int(version), int(D), int(K)))


def knn_points_backward(p1, p2, lengths1, lengths2, idxs, norm, grad_dists):
    """knn.h:127-134.  Returns (grad_p1, grad_p2)."""
    lib = _lib.load()
    p1 = _cuda_f32(p1, "p1")
    p2 = _cuda_f32(p2, "p2")
    _same_device(p1, p2, "p1 and p2")
    grad_dists = _cuda_f32(grad_dists, "grad_dists")
    lengths1 = _cuda_i64(lengths1, "lengths1", p1)
    lengths2 = _cuda_i64(lengths2, "lengths2", p1)
    idxs = _cuda_i64(idxs, "idxs", p1)
    if norm not in (1, 2):
        raise RuntimeError("Norm must be 1 or 2.")
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    K = idxs.shape[2]
    grad_p1 = torch.empty_like(p1)
    grad_p2 = torch.empty_like(p2)
    with _on_device(p1.device):
        ws = _ws(lib.pops_knn_backward_workspace_bytes(N, P2, D), p1.device)
        st = lib.pops_knn_points_backward_ws(p1.data_ptr(), p2.data_ptr(), lengths1.data_ptr(),
                                             lengths2.data_ptr(), idxs.data_ptr(),
                                             grad_dists.data_ptr(), N, P1, P2, D, K, int(norm),
                                             grad_p1.data_ptr(), grad_p2.data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream(p1))
    _lib.check(st, "knn_points_backward")
    return grad_p1, grad_p2


def ball_query(p1, p2, lengths1, lengths2, K, radius):
    """ball_query.h:62-68.  Returns (idx int64 padded -1, dists float32 padded 0)."""
    lib = _lib.load()
    p1 = _cuda_f32(p1, "p1")
    p2 = _cuda_f32(p2, "p2")
    _same_device(p1, p2, "p1 and p2")
    lengths1 = _cuda_i64(lengths1, "lengths1", p1)
    lengths2 = _cuda_i64(lengths2, "lengths2", p1)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    K = int(K)
    idx = torch.empty((N, P1, K), dtype=torch.int64, device=p1.device)
    dists = torch.empty((N, P1, K), dtype=torch.float32, device=p1.device)
    if idx.numel() == 0:
        return idx, dists
    if P2 == 0 or D == 0:
        return idx.fill_(-1), dists.zero_()
    with _on_device(p1.device):
        ws = _ws(lib.pops_ball_query_workspace_bytes(N, P1, P2, D, K), p1.device)
        st = lib.pops_ball_query(p1.data_ptr(), p2.data_ptr(), lengths1.data_ptr(),
                                 lengths2.data_ptr(), N, P1, P2, D, K, float(radius),
                                 idx.data_ptr(), dists.data_ptr(), ws.data_ptr(), ws.numel(),
                                 _stream(p1))
    _lib.check(st, "ball_query")
    return idx, dists


def sample_farthest_points(points, lengths, K, start_idxs, max_K=None):
    """sample_farthest_points.h:55-59.  `max_K` (optional, additive) avoids the device sync the
    reference pays for torch::max(K) (sample_farthest_points.cu:132)."""
    lib = _lib.load()
    points = _cuda_f32(points, "points")
    lengths = _cuda_i64(lengths, "lengths", points)
    K = _cuda_i64(K, "K", points)
    start_idxs = _cuda_i64(start_idxs, "start_idxs", points)
    N, P, D = points.shape
    if max_K is None:
        max_K = int(K.max().item()) if N > 0 else 0
    idx = torch.empty((N, max_K), dtype=torch.int64, device=points.device)
    if idx.numel() == 0:
        return idx
    if P == 0:
        idx.fill_(-1)
        idx[:, 0] = start_idxs
        return idx
    with _on_device(points.device):
        ws = _ws(lib.pops_fps_workspace_bytes(N, P, D, max_K), points.device)
        st = lib.pops_sample_farthest_points(points.data_ptr(), lengths.data_ptr(), K.data_ptr(),
                                             start_idxs.data_ptr(), N, P, D, max_K,
                                             idx.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _stream(points))
    _lib.check(st, "sample_farthest_points")
    return idx


def packed_to_padded(inputs_packed, first_idxs, max_size):
    """packed_to_padded_tensor.h:78-94: (F,D) -> (N,max_size,D), zero padded."""
    lib = _lib.load()
    x = _cuda_f32(inputs_packed, "inputs_packed")
    first = _cuda_i64(first_idxs, "first_idxs", x)
    if x.dim() != 2:
        raise RuntimeError("inputs_packed must be a 2-dimensional tensor")
    F, D = x.shape
    B = first.shape[0]
    out = torch.empty((B, int(max_size), D), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    with _on_device(x.device):
        st = lib.pops_packed_to_padded(x.data_ptr(), first.data_ptr(), F, B, int(max_size), D,
                                       out.data_ptr(), _stream(x))
    _lib.check(st, "packed_to_padded")
    return out


def padded_to_packed(inputs_padded, first_idxs, num_inputs):
    """packed_to_padded_tensor.h:97-113: (N,M,D) -> (F,D)."""
    lib = _lib.load()
    x = _cuda_f32(inputs_padded, "inputs_padded")
    first = _cuda_i64(first_idxs, "first_idxs", x)
    if x.dim() != 3:
        raise RuntimeError("inputs_padded must be a 3-dimensional tensor")
    B, M, D = x.shape
    out = torch.empty((int(num_inputs), D), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    with _on_device(x.device):
        st = lib.pops_padded_to_packed(x.data_ptr(), first.data_ptr(), int(num_inputs), B, M, D,
                                       out.data_ptr(), _stream(x))
    _lib.check(st, "padded_to_packed")
    return out


# ------------------------------------------------------------------------------------------------
# additive fused entry points
# ------------------------------------------------------------------------------------------------
def gather(x, idx, lengths, mode, oob_flag=None):
    """out[n,l,k,:] = x[n, idx[n,l,k], :] with knn_gather / masked_gather masking (one pass).

    x (N,M,U) f32, idx (N,L,K) i64, lengths (N) i64 or None -> (N,L,K,U)."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    idx = _cuda_i64(idx, "idx", x)
    if lengths is not None:
        lengths = _cuda_i64(lengths, "lengths", x)
    N, M, U = x.shape
    _, L, K = idx.shape
    out = torch.empty((N, L, K, U), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    if M == 0:
        return out.zero_()
    with _on_device(x.device):
        st = lib.pops_gather(x.data_ptr(), idx.data_ptr(), _ptr(lengths), N, M, U, L, K, int(mode),
                             out.data_ptr(), _ptr(oob_flag), _stream(x))
    _lib.check(st, "gather")
    return out


def gather_backward(grad_out, idx, lengths, M, mode):
    """Scatter-add of grad_out (N,L,K,U) into grad_x (N,M,U)."""
    lib = _lib.load()
    g = _cuda_f32(grad_out, "grad_out")
    idx = _cuda_i64(idx, "idx", g)
    if lengths is not None:
        lengths = _cuda_i64(lengths, "lengths", g)
    N, L, K, U = g.shape
    grad_x = torch.empty((N, int(M), U), dtype=torch.float32, device=g.device)
    if grad_x.numel() == 0:
        return grad_x
    with _on_device(g.device):
        st = lib.pops_gather_backward(g.data_ptr(), idx.data_ptr(), _ptr(lengths), N, int(M), U, L,
                                      K, int(mode), grad_x.data_ptr(), _stream(g))
    _lib.check(st, "gather_backward")
    return grad_x


RED = {None: 0, "sum": 1, "mean": 2, "max": 3}


def _ptr_array(tensors):
    import ctypes

    arr = (ctypes.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _chan_array(tensors):
    import ctypes

    arr = (ctypes.c_int64 * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.shape[2]
    return arr


def _check_feature_shapes(xfs, yfs, N, P1, P2):
    """The kernels index features with (N, P1) / (N, P2) of the points and ONE channel count per pair;
    anything else would read or write out of bounds (the reference fails in knn_gather /
    cosine_similarity broadcasting, functions/chamfer.py:143-160)."""
    if len(xfs) != len(yfs):
        raise ValueError("x and y must bring the same number of feature tensors")
    for f, (a, b) in enumerate(zip(xfs, yfs)):
        if a.dim() != 3 or b.dim() != 3:
            raise ValueError(f"feature {f}: expected tensors of shape (N, P, C)")
        if a.shape[0] != N or a.shape[1] != P1:
            raise ValueError(f"feature {f}: x feature has shape {tuple(a.shape)}, expected ({N}, {P1}, C)")
        if b.shape[0] != N or b.shape[1] != P2:
            raise ValueError(f"feature {f}: y feature has shape {tuple(b.shape)}, expected ({N}, {P2}, C)")
        if a.shape[2] != b.shape[2]:
            raise ValueError(f"feature {f}: x and y features differ in channels ({a.shape[2]} vs {b.shape[2]})")


def chamfer_forward(dists, idx, lengths1, lengths2, weights, P2, xfs, yfs, point_reduction, abs_cosine,
                    out=None):
    """Fused per-direction chamfer post-processing (see include/pointops_b200.h).
    dists/idx (N,P1) from the K=1 search; xfs/yfs lists of (N,P,C) feature tensors.
    out: optional (1+F, N) float32 buffer for the "sum" / "mean" reductions (row 0 = cham, rows
    1.. = features) so that two directions can land in one tensor.
    Returns (cham, feats (F,...) or None, argmax or None)."""
    lib = _lib.load()
    dists = _cuda_f32(dists, "dists")
    idx = _cuda_i64(idx, "idx", dists)
    N, P1 = dists.shape
    F = len(xfs)
    xfs = [_cuda_f32(t, "x_feature") for t in xfs]
    yfs = [_cuda_f32(t, "y_feature") for t in yfs]
    _check_feature_shapes(xfs, yfs, N, P1, int(P2))
    if idx.shape != dists.shape:
        raise ValueError("dists and idx must both be (N, P1)")
    lengths1 = _cuda_i64(lengths1, "lengths1", dists)
    lengths2 = _cuda_i64(lengths2, "lengths2", dists)
    red = RED[point_reduction]
    shape = (N, P1) if red == 0 else (N,)
    if out is not None:
        assert red in (1, 2) and out.shape == (1 + F, N) and out.is_contiguous() and out.dtype == torch.float32
        cham, feats = out[0], (out[1:] if F else None)
    else:
        cham = torch.empty(shape, dtype=torch.float32, device=dists.device)
        feats = torch.empty((F,) + shape, dtype=torch.float32, device=dists.device) if F else None
    argmax = torch.empty((N,), dtype=torch.int64, device=dists.device) if red == 3 else None
    if N == 0:
        return cham, feats, argmax
    if weights is not None:
        weights = _cuda_f32(weights, "weights")
    with _on_device(dists.device):
        st = lib.pops_chamfer_forward(dists.data_ptr(), idx.data_ptr(), lengths1.data_ptr(),
                                      lengths2.data_ptr(), _ptr(weights), N, P1, int(P2), F,
                                      _ptr_array(xfs), _ptr_array(yfs), _chan_array(xfs), red,
                                      int(bool(abs_cosine)), cham.data_ptr(), _ptr(feats), _ptr(argmax),
                                      _stream(dists))
    _lib.check(st, "chamfer_forward")
    return cham, feats, argmax


def chamfer_backward(x, y, idx, lengths1, lengths2, weights, norm, xfs, yfs, point_reduction, abs_cosine,
                     g_cham, g_feat, argmax, into=None, g_broadcast=False, g_scale=1.0):
    """Returns (grad_x, grad_y, [grad_xf...], [grad_yf...]).
    into: optional (grad_x, grad_y, [grad_xf...], [grad_yf...]) of already initialised buffers the
    call ADDS to (accumulate mode of pops_chamfer_backward).
    g_broadcast: g_cham is one scalar and g_feat one scalar per feature for all clouds, times g_scale."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    y = _cuda_f32(y, "y")
    N, P1, D = x.shape
    P2 = y.shape[1]
    F = len(xfs)
    xfs = [_cuda_f32(t, "x_feature") for t in xfs]
    yfs = [_cuda_f32(t, "y_feature") for t in yfs]
    _check_feature_shapes(xfs, yfs, N, P1, P2)
    idx = _cuda_i64(idx, "idx", x)
    if idx.numel() != N * P1:
        raise ValueError("idx must be (N, P1)")
    lengths1 = _cuda_i64(lengths1, "lengths1", x)
    lengths2 = _cuda_i64(lengths2, "lengths2", x)
    if into is not None:
        grad_x, grad_y, gxf, gyf = into
        for t, r in [(grad_x, x), (grad_y, y)] + list(zip(gxf, xfs)) + list(zip(gyf, yfs)):
            assert t.shape == r.shape and t.is_contiguous() and t.dtype == torch.float32 and t.device == r.device
    else:
        grad_x = torch.empty_like(x)
        grad_y = torch.empty_like(y)
        gxf = [torch.empty_like(t) for t in xfs]
        gyf = [torch.empty_like(t) for t in yfs]
    g_cham = _cuda_f32(g_cham, "g_cham")
    if g_feat is not None:
        g_feat = _cuda_f32(g_feat, "g_feat")
    with _on_device(x.device):
        st = lib.pops_chamfer_backward(x.data_ptr(), y.data_ptr(), idx.data_ptr(), lengths1.data_ptr(),
                                       lengths2.data_ptr(), _ptr(weights), N, P1, P2, D, int(norm), F,
                                       _ptr_array(xfs), _ptr_array(yfs), _chan_array(xfs),
                                       RED[point_reduction], int(bool(abs_cosine)), g_cham.data_ptr(),
                                       _ptr(g_feat), _ptr(argmax), grad_x.data_ptr(), grad_y.data_ptr(),
                                       _ptr_array(gxf), _ptr_array(gyf), int(into is not None), int(bool(g_broadcast)),
                                       float(g_scale), _stream(x))
    _lib.check(st, "chamfer_backward")
    return grad_x, grad_y, gxf, gyf


def sample_pdf(bins, weights, outputs, eps):
    """sample_pdf.h:58-78: bins (B,n_bins+1), weights (B,n_bins), outputs (B,n_samples) float32;
    `outputs` holds uniform numbers on entry and the samples on return (in place; its autograd
    version is bumped like torch::autograd::increment_version in sample_pdf_cpu.cpp:141)."""
    bins = _cuda_f32(bins, "bins")
    weights = _cuda_f32(weights, "weights")
    _same_device(bins, weights, "bins and weights")
    _same_device(bins, outputs, "bins and outputs")
    if not outputs.is_cuda or outputs.dtype != torch.float32:
        raise RuntimeError("outputs must be a float32 CUDA tensor.")
    if not outputs.is_contiguous():
        raise RuntimeError("outputs must be contiguous.")  # written in place: no silent copy
    if bins.ndim != 2 or weights.ndim != 2 or outputs.ndim != 2:
        raise RuntimeError("bins, weights and outputs must be 2-dimensional.")
    B, n_bins = weights.shape
    if bins.shape[0] != B or outputs.shape[0] != B:
        raise RuntimeError("Batch dimensions of bins, weights and outputs must agree.")
    if bins.shape[1] != n_bins + 1:
        raise RuntimeError("There must be one more bin edge than weights.")
    lib = _lib.load()
    with _on_device(bins.device):
        st = lib.pops_sample_pdf(bins.data_ptr(), weights.data_ptr(), outputs.data_ptr(), B, n_bins,
                                 outputs.shape[1], float(eps), _stream(bins))
    _lib.check(st, "sample_pdf")
    torch.autograd.graph.increment_version(outputs)
    return None


def point_covariances(x, idx, lengths):
    """Additive: neighbourhood gather + covariance in one kernel (functions/utils.py:111-153).
    x (N,M,D<=4) f32, idx (N,P,K) i64, lengths (N) i64 or None -> (cov (N,P,D,D), nn (N,P,K,D))."""
    lib = _lib.load()
    x = _cuda_f32(x, "points")
    idx = _cuda_i64(idx, "idx", x)
    if lengths is not None:
        lengths = _cuda_i64(lengths, "lengths", x)
    N, M, D = x.shape
    _, P, K = idx.shape
    nn = torch.empty((N, P, K, D), dtype=torch.float32, device=x.device)
    cov = torch.empty((N, P, D, D), dtype=torch.float32, device=x.device)
    if nn.numel() == 0:
        return cov, nn
    with _on_device(x.device):
        st = lib.pops_point_covariances(x.data_ptr(), idx.data_ptr(), _ptr(lengths), N, P, M, D, K,
                                        nn.data_ptr(), cov.data_ptr(), _stream(x))
    _lib.check(st, "point_covariances")
    return cov, nn
