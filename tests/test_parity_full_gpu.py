"""GPU parity on the WHOLE headline outputs (VERDICT r1 "next" #1): every query of the T shape
(uniform and ragged), full clouds of C4, >= 16 k queries of C5, each against the CPU oracle run on
all host cores -- a pruning / filtering search can only fail by MISSING a neighbour, which sampled
or property checks cannot see.  Plus the reference's naive FPS, reference-generated goldens for
get_point_covariances, and the defined behaviour on non-finite inputs.

Bar: indices and distances `torch.equal` (same unfused float32 arithmetic)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _threads():
    return max(1, os.cpu_count() or 1)


def _C():
    from pytorch3d_pointops_b200 import _C as C

    return C


# ------------------------------------------------------------------------------------------------
# T shape: B=32, P=16384, D=3, K=16 -- all 524 288 queries
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ragged", [False, True])
def test_t_shape_every_query_vs_oracle(oracle, ragged):
    gen = torch.Generator().manual_seed(0)
    N, P, K = 32, 16384, 16
    p = torch.rand(N, P, 3, generator=gen)  # bench.py's seed-0 batch
    lengths = torch.full((N,), P, dtype=torch.int64)
    if ragged:
        lengths = torch.randint(8192, P + 1, (N,), generator=gen)  # SURVEY 8(d) ragged variant
    pd, ld = p.to(DEV), lengths.to(DEV)
    idx, dists = _C().knn_points_idx(pd, pd, ld, ld, 2, K, -1)
    oi, od = oracle.knn_points_idx(p, p, lengths, lengths, 2, K, threads=_threads())
    assert torch.equal(idx.cpu(), oi)
    assert torch.equal(dists.cpu(), od)


def test_t_shape_host_pipeline_every_query(oracle):
    """host.HostKnn (the e2e path of bench.py: sliced search, pinned D2H) returns the same full result."""
    from pytorch3d_pointops_b200.host import HostKnn

    gen = torch.Generator().manual_seed(0)
    N, P, K = 32, 16384, 16
    p = torch.rand(N, P, 3, generator=gen)
    lengths = torch.randint(8192, P + 1, (N,), generator=gen)
    hk = HostKnn(N, P, P, 3, K, torch.device(DEV), slices=8)
    dists, idx = hk(p.pin_memory(), None, lengths.pin_memory())
    torch.cuda.synchronize()
    oi, od = oracle.knn_points_idx(p, p, lengths, lengths, 2, K, threads=_threads())
    assert torch.equal(idx, oi)
    assert torch.equal(dists, od)


def test_chamfer_search_every_query_vs_oracle(oracle):
    """C2's two K=1 searches (pair pre-pass) over all 32 ragged clouds of 8192 points, both directions."""
    gen = torch.Generator().manual_seed(1)
    N, P = 32, 8192
    x, y = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)
    xl = torch.randint(4096, P + 1, (N,), generator=gen)
    yl = torch.randint(4096, P + 1, (N,), generator=gen)
    i12, d12, i21, d21 = _C().knn_points_idx_pair(x.to(DEV), y.to(DEV), xl.to(DEV), yl.to(DEV), 2, 1)
    oi, od = oracle.knn_points_idx(x, y, xl, yl, 2, 1, threads=_threads())
    assert torch.equal(i12.cpu(), oi) and torch.equal(d12.cpu(), od)
    oi, od = oracle.knn_points_idx(y, x, yl, xl, 2, 1, threads=_threads())
    assert torch.equal(i21.cpu(), oi) and torch.equal(d21.cpu(), od)


# ------------------------------------------------------------------------------------------------
# C5: D=128, K=16, P=32768 -- two full clouds on the GPU, an 8192-query window of each vs the oracle
# ------------------------------------------------------------------------------------------------
def test_c5_two_clouds_16k_queries_vs_oracle(oracle):
    gen = torch.Generator().manual_seed(4)
    P, D, K = 32768, 128, 16
    x = torch.randn(2, P, D, generator=gen)
    L = torch.full((2,), P, dtype=torch.int64)
    idx, dists = _C().knn_points_idx(x.to(DEV), x.to(DEV), L.to(DEV), L.to(DEV), 2, K, -1)
    q0, q1 = 12000, 12000 + 8192
    oi, od = oracle.knn_points_idx(x, x, L, L, 2, K, q0=q0, q1=q1, threads=_threads())
    assert torch.equal(idx[:, q0:q1].cpu(), oi[:, q0:q1])
    assert torch.equal(dists[:, q0:q1].cpu(), od[:, q0:q1])


# ------------------------------------------------------------------------------------------------
# C4: ball_query K=32 r=0.1 on 32 full clouds of 16384 points (of the 128 of the config)
# ------------------------------------------------------------------------------------------------
def test_c4_full_clouds_vs_oracle(oracle):
    gen = torch.Generator().manual_seed(3)
    N, P, K, r = 32, 16384, 32, 0.1
    p = torch.rand(N, P, 3, generator=gen)
    L = torch.full((N,), P, dtype=torch.int64)
    idx, dists = _C().ball_query(p.to(DEV), p.to(DEV), L.to(DEV), L.to(DEV), K, r)
    oi, od = oracle.ball_query_idx(p, p, L, L, K=K, radius=r, threads=_threads())
    assert torch.equal(idx.cpu(), oi)
    assert torch.equal(dists.cpu(), od)


# ------------------------------------------------------------------------------------------------
# a9: the repo's own sample_farthest_points_naive
# ------------------------------------------------------------------------------------------------
def test_naive_fps_matches_reference_golden(golden):
    """fps_cases.npz `big.idx` was asserted equal to the reference's naive output when it was
    generated (tests/golden/make_golden.py: `assert torch.equal(si, si2)`)."""
    from pytorch3d_pointops_b200.functions.sample_farthest_points import (sample_farthest_points,
                                                                          sample_farthest_points_naive)

    g = golden("fps_cases")
    pts = g.t("big.points", DEV)
    sp, si = sample_farthest_points_naive(pts, K=200)
    assert torch.equal(si.cpu(), g.t("big.idx"))
    assert torch.equal(sp.cpu(), g.t("big.sampled"))
    sp2, si2 = sample_farthest_points(pts, K=200)
    assert torch.equal(si2, si) and torch.equal(sp2, sp)
    # ragged lengths, per-cloud K (list) and the -1 padding
    pts = g.t("ragged.points", DEV)
    sp, si = sample_farthest_points_naive(pts, g.t("ragged.lengths", DEV), g.t("ragged.K").tolist())
    assert torch.equal(si.cpu(), g.t("ragged.idx"))
    assert torch.equal(sp.cpu(), g.t("ragged.sampled"))


# ------------------------------------------------------------------------------------------------
# f2: get_point_covariances against outputs of the reference (functions/utils.py:111-153)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,K", [("cov3", 8), ("cov3", 16), ("cov2", 6)])
@pytest.mark.parametrize("with_grad", [False, True])
def test_point_covariances_reference_golden(golden, case, K, with_grad):
    from pytorch3d_pointops_b200.functions.utils import get_point_covariances

    g = golden("utils_cases")
    pts = g.t(f"{case}.points", DEV)
    if with_grad:  # the autograd formulation; without grad the fused gather + covariance kernel
        pts = pts.clone().requires_grad_(True)
    cov, nn = get_point_covariances(pts, g.t(f"{case}.lengths", DEV), K)
    assert torch.equal(nn.detach().cpu(), g.t(f"{case}.K{K}.nn"))
    want = g.t(f"{case}.K{K}.cov")
    assert torch.allclose(cov.detach().cpu(), want, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# non-finite inputs: defined, memory-safe behaviour (pytorch3d_pointops_b200/csrc/common.cuh)
# ------------------------------------------------------------------------------------------------
def _total_order_knn(p1, p2, L2, K, norm=2):
    """numpy restatement of the documented order: unfused float32 distance, finite < +inf < NaN,
    ties by lower index; (0, 0) padding beyond min(K, L2)."""
    a, b = p1.numpy().astype(np.float32), p2.numpy().astype(np.float32)[:L2]
    with np.errstate(all="ignore"):
        d = np.zeros((a.shape[0], b.shape[0]), np.float32)
        for c in range(a.shape[1]):
            diff = (a[:, None, c] - b[None, :, c]).astype(np.float32)
            d = (d + (np.abs(diff) if norm == 1 else (diff * diff).astype(np.float32))).astype(np.float32)
    idx = np.zeros((a.shape[0], K), np.int64)
    dist = np.zeros((a.shape[0], K), np.float32)
    for i in range(a.shape[0]):
        nan = np.isnan(d[i])
        order = np.lexsort((np.arange(b.shape[0]), np.where(nan, 0, d[i]), nan))[:K]
        idx[i, :len(order)] = order
        dist[i, :len(order)] = d[i][order]
    return torch.from_numpy(idx), torch.from_numpy(dist)


def _same_with_nan(got, want):
    gn, wn = torch.isnan(got), torch.isnan(want)
    return torch.equal(gn, wn) and torch.equal(got[~gn], want[~wn])


NONFINITE_SHAPES = [(3, 2000, 3, 16), (3, 300, 3, 5), (2, 400, 2, 4), (2, 300, 5, 7), (2, 1500, 3, 40),
                    (2, 1024, 64, 8)]


@pytest.mark.parametrize("N,P,D,K", NONFINITE_SHAPES)
def test_knn_nan_query_and_nan_point(N, P, D, K):
    """A NaN QUERY row returns the first K points with NaN distances (the reference's result: its push
    rule `size < K || dist < top` admits exactly the first K, knn_cpu.cpp:52); a NaN POINT ranks last;
    every other row and every other cloud is exact."""
    gen = torch.Generator().manual_seed(100 + D + K)
    p1, p2 = torch.rand(N, P, D, generator=gen), torch.rand(N, P, D, generator=gen)
    p1[0, 17, 1 % D] = float("nan")
    p2[1, 23, 0] = float("nan")
    L = torch.full((N,), P, dtype=torch.int64)
    L[N - 1] = P - 37
    idx, dists = _C().knn_points_idx(p1.to(DEV), p2.to(DEV), L.to(DEV), L.to(DEV), 2, K, -1)
    torch.cuda.synchronize()
    idx, dists = idx.cpu(), dists.cpu()
    assert (idx >= 0).all() and (idx < P).all()
    for n in range(N):
        ln = int(L[n])
        wi, wd = _total_order_knn(p1[n, :ln], p2[n], ln, K)
        assert torch.equal(idx[n, :ln], wi), n
        assert _same_with_nan(dists[n, :ln], wd), n
        assert not idx[n, ln:].any() and not dists[n, ln:].any()
    assert torch.equal(idx[0, 17], torch.arange(K)) and torch.isnan(dists[0, 17]).all()


@pytest.mark.parametrize("N,P,D,K", NONFINITE_SHAPES[:5])
def test_knn_inf_and_huge_coordinates_match_oracle(oracle, N, P, D, K):
    """+inf distances are ordinary values for the reference's heap, and so are coordinates whose squares
    overflow: the result is well defined and must equal the oracle's bit for bit."""
    gen = torch.Generator().manual_seed(200 + D + K)
    p1, p2 = torch.rand(N, P, D, generator=gen), torch.rand(N, P, D, generator=gen)
    p2[0, 2, 0] = float("inf")
    p2[1, :, :] *= 3e19  # squared distances overflow to +inf for most pairs
    p1[1, :, :] *= 3e19
    L1 = torch.full((N,), P, dtype=torch.int64)
    L2 = torch.full((N,), P, dtype=torch.int64)
    L2[0] = min(P, K)  # every point of the cloud is in every list: the +inf one last
    idx, dists = _C().knn_points_idx(p1.to(DEV), p2.to(DEV), L1.to(DEV), L2.to(DEV), 2, K, -1)
    torch.cuda.synchronize()
    oi, od = oracle.knn_points_idx(p1, p2, L1, L2, 2, K, threads=_threads())
    assert not torch.isnan(od).any()
    assert torch.equal(idx.cpu(), oi)
    assert torch.equal(dists.cpu(), od)


def test_knn_pair_and_sliced_paths_with_nan():
    C = _C()
    gen = torch.Generator().manual_seed(7)
    N, P = 3, 2048
    x, y = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)
    x[1, 100, 2] = float("nan")
    L = torch.full((N,), P, dtype=torch.int64)
    xd, yd, Ld = x.to(DEV), y.to(DEV), L.to(DEV)
    i12, d12, i21, d21 = C.knn_points_idx_pair(xd, yd, Ld, Ld, 2, 1)
    a_i, a_d = C.knn_points_idx(xd, yd, Ld, Ld, 2, 1, -1)
    b_i, b_d = C.knn_points_idx(yd, xd, Ld, Ld, 2, 1, -1)
    torch.cuda.synchronize()
    assert torch.equal(i12, a_i) and _same_with_nan(d12.cpu(), a_d.cpu())
    assert torch.equal(i21, b_i) and _same_with_nan(d21.cpu(), b_d.cpu())
    for n in range(N):
        wi, wd = _total_order_knn(x[n], y[n], P, 1)
        assert torch.equal(i12[n].cpu(), wi) and _same_with_nan(d12[n].cpu(), wd)
        wi, wd = _total_order_knn(y[n], x[n], P, 1)
        assert torch.equal(i21[n].cpu(), wi) and _same_with_nan(d21[n].cpu(), wd)
    ks = C.KnnSliced(xd, yd, Ld, Ld, 2, 4)
    ks.prepare()
    ks.search(0, 2)
    ks.search(2, 3)
    full_i, full_d = C.knn_points_idx(xd, yd, Ld, Ld, 2, 4, -1)
    torch.cuda.synchronize()
    assert torch.equal(ks.idx, full_i) and _same_with_nan(ks.dists.cpu(), full_d.cpu())


def test_chamfer_with_nan_point_is_nan_not_a_fault():
    """one NaN point used to leave idx = 0xFFFFFFFF in the K=1 lists, which the
    chamfer kernels dereferenced.  The reference returns a NaN loss; so do we, and the device stays
    healthy."""
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    gen = torch.Generator().manual_seed(9)
    N, P = 4, 3000
    x = torch.rand(N, P, 3, generator=gen)
    y = torch.rand(N, P, 3, generator=gen)
    x[2, 11, 0] = float("nan")
    xn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=gen), dim=-1)
    yn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=gen), dim=-1)
    for batch_reduction in ("mean", None):
        xd = x.to(DEV).requires_grad_(True)
        yd = y.to(DEV).requires_grad_(True)
        loss, lf = chamfer_distance(xd, yd, x_features={"normals": xn.to(DEV)}, y_features={"normals": yn.to(DEV)},
                                    feature_names=["normals"], batch_reduction=batch_reduction)
        (loss.sum() + lf["normals"].sum()).backward()
        torch.cuda.synchronize()
        if batch_reduction is None:
            assert torch.isnan(loss[2]) and torch.isfinite(loss[[0, 1, 3]]).all()
            assert torch.isfinite(xd.grad[[0, 1, 3]]).all() and torch.isfinite(yd.grad[[0, 1, 3]]).all()
        else:
            assert torch.isnan(loss)
    # single direction, weights, max reduction: the per-direction autograd node
    loss, _ = chamfer_distance(x.to(DEV), y.to(DEV), single_directional=True, point_reduction="max",
                               batch_reduction=None, weights=torch.ones(N, device=DEV))
    torch.cuda.synchronize()
    assert torch.isfinite(loss[[0, 1, 3]]).all()


def test_ball_query_non_finite_matches_oracle(oracle):
    """NaN never satisfies d2 < r2 (ball_query_cpu.cpp:44); +inf and huge coordinates neither -- the
    finite points of the same cloud must still be found."""
    gen = torch.Generator().manual_seed(12)
    N, P, K, r = 3, 3000, 8, 0.15
    p1, p2 = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)
    p2[0, 40, 1] = float("nan")
    p2[0, 41, 0] = float("inf")
    p1[0, 7, 2] = float("nan")
    p2[1, 5, :] = 3e19
    p1[2, 9, 0] = float("inf")
    L = torch.full((N,), P, dtype=torch.int64)
    for P1 in (P, 600):  # buffered scan kernel (P1 >= 1024) and the thread-per-query kernel
        idx, dists = _C().ball_query(p1[:, :P1].contiguous().to(DEV), p2.to(DEV),
                                     torch.full((N,), P1, dtype=torch.int64, device=DEV), L.to(DEV), K, r)
        torch.cuda.synchronize()
        oi, od = oracle.ball_query_idx(p1[:, :P1].contiguous(), p2, torch.full((N,), P1, dtype=torch.int64), L,
                                       K=K, radius=r, threads=_threads())
        assert torch.equal(idx.cpu(), oi)
        assert torch.equal(dists.cpu(), od)


# ------------------------------------------------------------------------------------------------
# the rewritten HBM kernels: odd alignments, every branch
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("U", [1, 2, 3, 5, 8, 12])
@pytest.mark.parametrize("mode", ["knn", "masked"])
def test_gather_stream_kernel_vs_torch(U, mode):
    from pytorch3d_pointops_b200.functions.knn import knn_gather
    from pytorch3d_pointops_b200.functions.utils import masked_gather

    gen = torch.Generator().manual_seed(300 + U)
    N, M, L, K = 5, 257, 131, 7  # L*K*U odd for odd U: clouds start at every 4-byte misalignment
    x = torch.rand(N, M, U, generator=gen).to(DEV)
    idx = torch.randint(0, M, (N, L, K), generator=gen).to(DEV)
    if mode == "knn":
        lengths = torch.tensor([M, 3, 7, 0, 100], device=DEV)  # lengths < K: trailing slots zero
        out = knn_gather(x, idx, lengths)
        want = x[torch.arange(N, device=DEV)[:, None, None], idx]
        mask = torch.arange(K, device=DEV)[None, None, :] >= lengths[:, None, None]
        want = torch.where(mask[..., None], torch.zeros_like(want), want)
        assert torch.equal(out, want)
        assert torch.equal(knn_gather(x, idx), x[torch.arange(N, device=DEV)[:, None, None], idx])
        bad = idx.clone()
        bad[3, 5, 2] = -1
        with pytest.raises(RuntimeError, match="out of bounds"):
            knn_gather(x, bad)
    else:
        idx[torch.rand(N, L, K, generator=gen).to(DEV) < 0.3] = -1
        out = masked_gather(x, idx)
        want = x[torch.arange(N, device=DEV)[:, None, None], idx.clamp(min=0)]
        want = torch.where((idx < 0)[..., None], torch.zeros_like(want), want)
        assert torch.equal(out, want)


@pytest.mark.parametrize("D", [1, 3, 5, 8])
def test_packed_padded_segment_kernels_vs_oracle(oracle, D):
    from pytorch3d_pointops_b200.functions.packed_to_padded import packed_to_padded, padded_to_packed

    gen = torch.Generator().manual_seed(400 + D)
    lens = torch.tensor([0, 17, 1, 333, 0, 64, 129, 2])
    first = torch.cumsum(lens, 0) - lens
    F, max_size = int(lens.sum()), int(lens.max())
    packed = torch.rand(F, D, generator=gen)
    padded = packed_to_padded(packed.to(DEV), first.to(DEV), max_size)
    want = oracle.packed_to_padded_C(packed, first, max_size)
    assert torch.equal(padded.cpu(), want)
    back = padded_to_packed(padded, first.to(DEV), F)
    assert torch.equal(back.cpu(), packed)
    assert torch.equal(back.cpu(), oracle.padded_to_packed_C(want, first, F))
    # rows in front of the first cloud and a max_size below the longest cloud
    first2 = first + 5
    got = _C().padded_to_packed(padded, first2.to(DEV), F + 5)
    assert torch.equal(got.cpu(), oracle.padded_to_packed_C(want, first2, F + 5))
    small = _C().packed_to_padded(packed.to(DEV), first.to(DEV), 100)
    assert torch.equal(small.cpu(), want[:, :100])
    # 1-D inputs of the public API
    v = torch.rand(F, generator=gen)
    assert torch.equal(packed_to_padded(v.to(DEV), first.to(DEV), max_size).cpu(),
                       oracle.packed_to_padded_C(v[:, None], first, max_size)[..., 0])


@pytest.mark.parametrize("D", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("norm", [1, 2])
@pytest.mark.parametrize("K", [1, 5, 16])
def test_knn_backward_rows_kernel_vs_oracle(oracle, D, norm, K):
    gen = torch.Generator().manual_seed(500 + 10 * D + K)
    N, P1, P2 = 3, 777, 530
    p1, p2 = torch.rand(N, P1, D, generator=gen), torch.rand(N, P2, D, generator=gen)
    l1, l2 = torch.tensor([P1, 100, 0]), torch.tensor([P2, 3, 77])
    idx = torch.randint(0, P2, (N, P1, K), generator=gen)
    idx[torch.rand(N, P1, K, generator=gen) < 0.1] = -1  # ball-query padding is skipped
    g = torch.randn(N, P1, K, generator=gen)
    g1, g2 = _C().knn_points_backward(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), idx.to(DEV), norm, g.to(DEV))
    w1, w2 = oracle.knn_points_backward(p1, p2, l1, l2, idx, norm, g)
    assert torch.equal(g1.cpu(), w1)  # summed in k order per (row, d): bit-exact
    assert torch.allclose(g2.cpu(), w2, rtol=1e-5, atol=1e-5 * float(w2.abs().max() + 1e-30))
