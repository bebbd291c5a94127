"""GPU parity of the tensor-core KNN path (32 <= D <= 256, L2, K <= 16: tcgen05 filter + exact
re-rank, csrc/knn_tc.cu) against the CPU oracle.  Indices AND distances bit-exact: the tensor
cores only select candidates, membership and order come from the reference's own arithmetic."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _C():
    from pytorch3d_pointops_b200 import _C as C

    return C


def _check(oracle, p1, p2, l1, l2, K):
    oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, K)
    gi, gd = _C().knn_points_idx(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), 2, K, -1)
    torch.cuda.synchronize()
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)


SWEEP = [
    # N, P1, P2, D, K, ragged
    (1, 128, 512, 128, 16, False), (2, 300, 1000, 128, 16, False), (2, 300, 1000, 64, 8, True),
    (3, 257, 2049, 32, 1, False), (2, 200, 1500, 100, 5, True), (2, 200, 1500, 256, 16, False),
    (2, 1000, 5000, 128, 16, True), (1, 64, 513, 36, 3, False), (2, 700, 640, 48, 16, True),
]


@pytest.mark.parametrize("N,P1,P2,D,K,ragged", SWEEP)
def test_oracle_sweep(oracle, N, P1, P2, D, K, ragged):
    gen = torch.Generator().manual_seed(N + P1 + P2 + D + K)
    p1 = torch.randn(N, P1, D, generator=gen)
    p2 = torch.randn(N, P2, D, generator=gen)
    l1 = torch.full((N,), P1, dtype=torch.int64)
    l2 = torch.full((N,), P2, dtype=torch.int64)
    if ragged:
        l1 = torch.randint(1, P1 + 1, (N,), generator=gen)
        l2 = torch.randint(1, P2 + 1, (N,), generator=gen)
        l2[0] = 7  # fewer points than K=8/16: (0, 0) padding
        l1[-1] = P1
    _check(oracle, p1, p2, l1, l2, K)


def test_massive_ties_take_the_exact_fallback(oracle):
    """Integer grids: hundreds of points at the same distance.  The approximate band cannot be
    certified, the queries are flagged and recomputed by knn_exact_rows_kernel."""
    gen = torch.Generator().manual_seed(5)
    p1 = torch.randint(0, 3, (1, 130, 128), generator=gen).float()
    p2 = torch.randint(0, 3, (1, 700, 128), generator=gen).float()
    L1, L2 = torch.tensor([130]), torch.tensor([700])
    _check(oracle, p1, p2, L1, L2, 16)
    # low-dimensional structure embedded in 64-D: many exact duplicates of each distance
    q = torch.randint(0, 2, (2, 300, 64), generator=gen).float()
    q[..., 6:] = 0
    _check(oracle, q, q.flip(1).contiguous(), torch.tensor([300, 211]), torch.tensor([300, 300]), 8)


def test_all_points_identical_takes_the_dense_fallback(oracle):
    """More flagged queries than knn_exact_rows_kernel accepts (4096): the gated generic kernel
    recomputes them.  Ties -> lowest indices."""
    p1 = torch.full((1, 4500, 32), 0.25)
    p2 = torch.full((1, 600, 32), 0.25)
    _check(oracle, p1, p2, torch.tensor([4500]), torch.tensor([600]), 4)


def test_adversarial_orders_scales_offsets(oracle):
    gen = torch.Generator().manual_seed(11)
    D = 64
    u = torch.nn.functional.normalize(torch.randn(D, generator=gen), dim=0)
    P2 = 3000
    # every point is closer than all previous ones: each is a new record for every query
    p2 = ((P2 - torch.arange(P2, dtype=torch.float32))[:, None] * u[None] * 0.01)[None].contiguous()
    p2 = p2 + 1e-3 * torch.randn(1, P2, D, generator=gen)
    p1 = 1e-2 * torch.randn(1, 200, D, generator=gen)
    _check(oracle, p1, p2, torch.tensor([200]), torch.tensor([P2]), 16)
    base1 = torch.randn(2, 150, D, generator=gen)
    base2 = torch.randn(2, 900, D, generator=gen)
    L1, L2 = torch.tensor([150, 99]), torch.tensor([900, 640])
    for f in (lambda t: t + 100.0, lambda t: t * 1e-12, lambda t: t * 1e12, lambda t: t.abs() + 3.0):
        _check(oracle, f(base1).contiguous(), f(base2).contiguous(), L1, L2, 16)


def test_python_api_forward_backward_highdim(oracle):
    """knn_points(return_nn=True) + backward on a shape that takes the tensor-core path."""
    from pytorch3d_pointops_b200.functions import knn_points

    gen = torch.Generator().manual_seed(2)
    p1 = torch.randn(2, 257, 128, generator=gen)
    p2 = torch.randn(2, 1100, 128, generator=gen)
    l1, l2 = torch.tensor([257, 100]), torch.tensor([1100, 777])
    a = p1.to(DEV).requires_grad_(True)
    b = p2.to(DEV).requires_grad_(True)
    res = knn_points(a, b, l1.to(DEV), l2.to(DEV), K=8, return_nn=True)
    oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, 8)
    assert torch.equal(res.idx.cpu(), oi) and torch.equal(res.dists.detach().cpu(), od)
    g = torch.randn(2, 257, 8, generator=gen)
    (res.dists * g.to(DEV)).sum().backward()
    o1, o2 = oracle.knn_points_backward(p1, p2, l1, l2, oi, 2, g)
    assert torch.equal(a.grad.cpu(), o1)
    assert torch.allclose(b.grad.cpu(), o2, rtol=1e-5, atol=1e-5 * float(o2.abs().max()))


def test_config5_shape_properties(oracle):
    """BASELINE.json configs[4] (D=128, K=16, P=32768; 2 of the 16 clouds to bound the run time):
    size-independent properties on the full output and exact oracle agreement on sampled queries."""
    gen = torch.Generator().manual_seed(4)
    N, P, D, K = 2, 32768, 128, 16
    x = torch.randn(N, P, D, generator=gen)
    L = torch.tensor([P, 30001])
    xd, Ld = x.to(DEV), L.to(DEV)
    idx, dists = _C().knn_points_idx(xd, xd, Ld, Ld, 2, K, -1)
    valid = torch.arange(P, device=DEV)[None] < Ld[:, None]
    assert torch.equal(idx[..., 0][valid], torch.arange(P, device=DEV)[None].expand(N, -1)[valid])
    assert not dists[..., 0].any()
    assert (dists[..., 1:] >= dists[..., :-1]).all()
    assert (idx < Ld[:, None, None]).all() and (idx >= 0).all()
    assert not idx[~valid].any() and not dists[~valid].any()
    for n, q0 in ((0, 0), (0, 20000), (1, 29990)):
        q1 = min(q0 + 24, int(L[n]))
        oi, od = oracle.knn_points_idx(x[n:n + 1], x[n:n + 1], L[n:n + 1], L[n:n + 1], 2, K, q0=q0, q1=q1, threads=8)
        assert torch.equal(idx[n, q0:q1].cpu(), oi[0, q0:q1])
        assert torch.equal(dists[n, q0:q1].cpu(), od[0, q0:q1])
