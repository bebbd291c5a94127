"""GPU parity: knn_points / knn_gather / backward against the reference's golden vectors and the
CPU oracle on seeded inputs.  Indices bit-exact (ties -> lower index), distances bit-exact
(same unfused float32 arithmetic), gradients rtol 1e-5 (float atomics reorder the sums)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

KNN_CASES = ["d3k1", "d3k16", "d3k32_l1", "d3k40", "d2k5", "d5k7", "d33k4", "d128k16", "klen"]


def _ops():
    from pytorch3d_pointops_b200 import _C
    from pytorch3d_pointops_b200.functions import knn_gather, knn_points

    return _C, knn_points, knn_gather


@pytest.mark.parametrize("name", KNN_CASES)
def test_golden_forward_backward(golden, name):
    _C, knn_points, _ = _ops()
    g = golden("knn_cases")
    p1 = g.t(f"{name}.p1", DEV).requires_grad_(True)
    p2 = g.t(f"{name}.p2", DEV).requires_grad_(True)
    l1, l2 = g.t(f"{name}.l1", DEV), g.t(f"{name}.l2", DEV)
    K, norm = int(g.a(f"{name}.K")), int(g.a(f"{name}.norm"))
    res = knn_points(p1, p2, l1, l2, norm=norm, K=K, return_nn=True)
    assert res.idx.dtype == torch.int64
    assert torch.equal(res.idx.cpu(), g.t(f"{name}.idx"))
    assert torch.equal(res.dists.detach().cpu(), g.t(f"{name}.dists"))
    assert torch.equal(res.knn.detach().cpu(), g.t(f"{name}.knn"))
    ((res.dists * g.t(f"{name}.gd", DEV)).sum() + (res.knn * g.t(f"{name}.gn", DEV)).sum()).backward()
    # float atomics reorder the sums: 1e-5 relative to the gradient scale
    for got, want in ((p1.grad.cpu(), g.t(f"{name}.grad_p1")), (p2.grad.cpu(), g.t(f"{name}.grad_p2"))):
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


def test_readme_config(golden):
    """BASELINE.json configs[0]: README Pointclouds (1000/800 pts), self-KNN K=8."""
    from pytorch3d_pointops_b200.structures import Pointclouds

    _, knn_points, _ = _ops()
    g = golden("knn_readme")
    pc = Pointclouds([g.t("p0", DEV), g.t("p1", DEV)])
    X, L = pc.points_padded(), pc.num_points_per_cloud()
    assert torch.equal(X.cpu(), g.t("padded"))
    out = knn_points(X, X, lengths1=L, lengths2=L, K=8)
    assert torch.equal(out.idx.cpu(), g.t("idx"))
    assert torch.equal(out.dists.cpu(), g.t("dists"))


@pytest.mark.parametrize("K", [1, 3, 7, 16, 17, 32, 33])
def test_exact_ties_integer_grid(golden, K):
    _C, _, _ = _ops()
    g = golden("knn_cases")
    p = g.t("grid.p", DEV)
    L = torch.tensor([p.shape[1]], device=DEV)
    idx, dists = _C.knn_points_idx(p, p, L, L, 2, K, -1)
    assert torch.equal(idx.cpu(), g.t(f"grid.K{K}.idx"))
    assert torch.equal(dists.cpu(), g.t(f"grid.K{K}.dists"))


def test_tie_vector(golden):
    _, knn_points, _ = _ops()
    g = golden("knn_cases")
    for K in (3, 4, 5):
        r = knn_points(g.t("tie.p1", DEV), g.t("tie.p2", DEV), K=K)
        assert torch.equal(r.idx.cpu(), g.t(f"tie.K{K}.idx"))


SWEEP = [
    # N, P1, P2, D, K, norm
    (2, 700, 900, 3, 16, 2), (3, 513, 2049, 3, 1, 2), (2, 100, 5000, 3, 32, 2),
    (1, 2000, 2000, 3, 8, 1), (2, 257, 300, 2, 5, 2), (2, 300, 257, 4, 12, 1),
    (2, 300, 257, 1, 3, 2), (1, 64, 200, 8, 8, 2), (2, 50, 120, 16, 40, 2),
    (1, 40, 300, 64, 4, 1), (1, 33, 97, 3, 100, 2), (2, 1, 1, 3, 1, 2), (1, 5, 3, 3, 7, 2),
    (1, 130, 70, 3, 64, 2), (1, 60, 500, 5, 200, 2),
]


@pytest.fixture
def force_tiled():
    """Keep D=3 L2 calls on the tiled brute-force kernel (the default route only for P2 < 64 and K > 32)."""
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    lib.pops_set_option(b"knn_order", 0)
    yield
    lib.pops_set_option(b"knn_order", -1)


@pytest.mark.parametrize("N,P1,P2,K", [(2, 700, 900, 16), (3, 513, 2049, 1), (2, 100, 5000, 32), (1, 300, 3000, 4)])
def test_tiled_d3_kernel_vs_oracle(oracle, force_tiled, N, P1, P2, K):
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(N * 77 + P1 + P2 + K)
    p1 = torch.randn(N, P1, 3, generator=gen)
    p2 = torch.randn(N, P2, 3, generator=gen)
    l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
    l2 = torch.randint(0, P2 + 1, (N,), generator=gen)
    l1[0], l2[0] = P1, P2
    oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, K)
    gi, gd = _C.knn_points_idx(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), 2, K, -1)
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)


@pytest.mark.parametrize("N,P1,P2,D,K,norm", SWEEP)
def test_oracle_sweep(oracle, N, P1, P2, D, K, norm):
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(N * 1000 + P1 + P2 + D + K + norm)
    p1 = torch.randn(N, P1, D, generator=gen)
    p2 = torch.randn(N, P2, D, generator=gen)
    l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
    l2 = torch.randint(0, P2 + 1, (N,), generator=gen)
    l1[0], l2[0] = P1, P2
    oi, od = oracle.knn_points_idx(p1, p2, l1, l2, norm, K)
    gi, gd = _C.knn_points_idx(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), norm, K, -1)
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)
    grad = torch.randn(N, P1, K, generator=gen)
    o1, o2 = oracle.knn_points_backward(p1, p2, l1, l2, oi, norm, grad)
    g1, g2 = _C.knn_points_backward(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), gi, norm, grad.to(DEV))
    assert torch.equal(g1.cpu(), o1)  # no atomics on p1: same op order as the reference
    assert torch.allclose(g2.cpu(), o2, rtol=1e-5, atol=1e-5 * max(1.0, float(o2.abs().max())))


def test_duplicates_offsets_and_scales(oracle):
    """Adversarial inputs for the filter: duplicated points, large offsets (cancellation in the
    expanded form), tiny and huge scales, clustered queries."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(77)
    base = torch.rand(2, 600, 3, generator=gen)
    cases = {
        "dups": base[:, torch.randint(0, 40, (600,), generator=gen)],
        "offset1e3": base + 1000.0,
        "offset1e5": base * 0.01 + 1e5,
        "tiny": base * 1e-20,
        "huge": base * 1e15,
        "mixed": torch.cat([base[:, :300] * 1e-3, base[:, 300:] * 50 + 7], 1),
        "line": torch.stack([base[..., 0], base[..., 0] * 0, base[..., 0] * 0], -1),
    }
    for name, p in cases.items():
        p = p.contiguous()
        L = torch.tensor([600, 431])
        for K in (1, 16):
            oi, od = oracle.knn_points_idx(p, p, L, L, 2, K)
            gi, gd = _C.knn_points_idx(p.to(DEV), p.to(DEV), L.to(DEV), L.to(DEV), 2, K, -1)
            assert torch.equal(gi.cpu(), oi), (name, K)
            assert torch.equal(gd.cpu(), od), (name, K)


def test_backward_many_to_one(oracle):
    """Every query shares one neighbour: heavy atomic contention on a single grad_p2 row."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(3)
    p1 = torch.rand(1, 4096, 3, generator=gen) + 5.0
    p2 = torch.cat([torch.full((1, 1, 3), 5.5), torch.rand(1, 63, 3, generator=gen) - 100.0], 1)
    oi, od = oracle.knn_points_idx(p1, p2, None, None, 2, 1)
    assert not oi.any()
    g = torch.rand(1, 4096, 1, generator=gen)
    o1, o2 = oracle.knn_points_backward(p1, p2, None, None, oi, 2, g)
    L1 = torch.tensor([4096], device=DEV)
    L2 = torch.tensor([64], device=DEV)
    g1, g2 = _C.knn_points_backward(p1.to(DEV), p2.to(DEV), L1, L2, oi.to(DEV), 2, g.to(DEV))
    assert torch.equal(g1.cpu(), o1)
    assert torch.allclose(g2.cpu(), o2, rtol=1e-5, atol=1e-4)


def test_knn_gather_semantics(golden):
    _, _, knn_gather = _ops()
    from pytorch3d_pointops_b200.functions import masked_gather

    g = golden("gather_cases")
    x = g.t("kg.x", DEV)
    idx = g.t("kg.idx", DEV)
    assert torch.equal(knn_gather(x, idx, g.t("kg.lengths", DEV)).cpu(), g.t("kg.out"))
    assert torch.equal(knn_gather(x, idx).cpu(), g.t("kg.out_full"))
    assert torch.equal(masked_gather(x, g.t("mg.idx3", DEV)).cpu(), g.t("mg.out3"))
    assert torch.equal(masked_gather(x, g.t("mg.idx2", DEV)).cpu(), g.t("mg.out2"))
    # -1 is an error for knn_gather in the reference
    with pytest.raises(RuntimeError, match="out of bounds"):
        knn_gather(x, g.t("mg.idx3", DEV))
    # vectorised path (U % 4 == 0) and other dtypes
    x8 = torch.randn(3, 40, 8, device=DEV)
    want = x8.cpu()[:, :, None].expand(-1, -1, 6, -1).gather(1, idx.cpu()[..., None].expand(-1, -1, -1, 8))
    assert torch.equal(knn_gather(x8, idx).cpu(), want)
    xi = torch.randint(0, 1 << 40, (3, 40, 3), device=DEV)
    wanti = xi.cpu()[:, :, None].expand(-1, -1, 6, -1).gather(1, idx.cpu()[..., None].expand(-1, -1, -1, 3))
    assert torch.equal(knn_gather(xi, idx).cpu(), wanti)
    # backward = scatter-add
    xr = x.clone().requires_grad_(True)
    w = torch.randn(3, 25, 6, 7, device=DEV)
    (knn_gather(xr, idx) * w).sum().backward()
    ref = torch.zeros_like(x).index_put_(
        (torch.arange(3, device=DEV)[:, None, None].expand_as(idx), idx), w, accumulate=True)
    assert torch.allclose(xr.grad, ref, rtol=1e-5, atol=1e-5)


def test_headline_shape_properties(oracle):
    """BASELINE target shape (B=32, P=16384, K=16, D=3): size-independent properties on the full
    output plus exact oracle agreement on a sample of queries."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(0)
    N, P, K = 32, 16384, 16
    p = torch.rand(N, P, 3, generator=gen)
    lengths = torch.randint(8192, P + 1, (N,), generator=gen)
    lengths[0] = P
    pd, ld = p.to(DEV), lengths.to(DEV)
    idx, dists = _C.knn_points_idx(pd, pd, ld, ld, 2, K, -1)
    valid = torch.arange(P, device=DEV)[None] < ld[:, None]
    # self is the nearest neighbour at distance 0 (random floats: no duplicates)
    assert torch.equal(idx[..., 0][valid], torch.arange(P, device=DEV)[None].expand(N, -1)[valid])
    assert not dists[..., 0].any()
    # ascending, in range, padding rows zero
    assert (dists[..., 1:] >= dists[..., :-1]).all()
    assert (idx < ld[:, None, None]).all() and (idx >= 0).all()
    assert not idx[~valid].any() and not dists[~valid].any()
    # distances recomputed from the indices agree bit for bit (unfused arithmetic)
    nb = pd[torch.arange(N, device=DEV)[:, None, None], idx]
    diff = pd[:, :, None, :] - nb
    sq = diff * diff
    recomputed = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    assert torch.equal(recomputed[valid], dists[valid])
    # exact agreement with the oracle on 64 queries of 3 clouds
    for n in (0, 7, 31):
        oi, od = oracle.knn_points_idx(p[n:n + 1], p[n:n + 1], lengths[n:n + 1], lengths[n:n + 1], 2, K,
                                       q0=1000, q1=1064, threads=8)
        assert torch.equal(idx[n, 1000:1064].cpu(), oi[0, 1000:1064])
        assert torch.equal(dists[n, 1000:1064].cpu(), od[0, 1000:1064])


def test_no_cuda_errors_and_streams():
    """Runs on a non-default stream and leaves no sticky CUDA error behind."""
    _C, knn_points, _ = _ops()
    s = torch.cuda.Stream(device=DEV)
    p = torch.rand(2, 3000, 3, device=DEV)
    ref = knn_points(p, p, K=4)
    s.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(s):
        out = knn_points(p, p, K=4)
    s.synchronize()
    assert torch.equal(out.idx, ref.idx) and torch.equal(out.dists, ref.dists)
    torch.cuda.synchronize()


def test_host_pipeline_matches_device_call():
    """host.HostKnn (one pre-pass, sliced search, copies on side streams) returns exactly what one
    knn_points_idx call returns."""
    from pytorch3d_pointops_b200.host import HostKnn

    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(9)
    N, P, K = 7, 3000, 8
    p = torch.rand(N, P, 3, generator=gen).pin_memory()
    L = torch.randint(1, P + 1, (N,), generator=gen).pin_memory()
    hk = HostKnn(N, P, P, 3, K, DEV, slices=3)
    for _ in range(2):  # second call reuses the staging
        d, i = hk(p, None, L)
        torch.cuda.synchronize()
        ri, rd = _C.knn_points_idx(p.to(DEV), p.to(DEV), L.to(DEV), L.to(DEV), 2, K, -1)
        assert torch.equal(i, ri.cpu()) and torch.equal(d, rd.cpu())
    q = torch.rand(N, 500, 3, generator=gen).pin_memory()
    hk2 = HostKnn(N, 500, P, 3, K, DEV, slices=4)
    d, i = hk2(q, p, None, L)
    torch.cuda.synchronize()
    L1 = torch.full((N,), 500, device=DEV)
    ri, rd = _C.knn_points_idx(q.to(DEV), p.to(DEV), L1, L.to(DEV), 2, K, -1)
    assert torch.equal(i, ri.cpu()) and torch.equal(d, rd.cpu())


def test_two_phase_range_api_matches_single_call():
    """pops_knn_points_prepare + pops_knn_points_idx_range, any partition of the batch, equals one
    pops_knn_points_idx call -- on the path with a shared pre-pass (D=3, L2, large P2) and on the
    paths where a range is simply a batch of its own (D=5; L1; tiny P2)."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(21)
    for (N, P1, P2, D, K, norm) in [(9, 700, 2500, 3, 8, 2), (5, 300, 400, 5, 6, 2), (4, 300, 1500, 3, 5, 1),
                                    (6, 100, 50, 3, 3, 2), (3, 1100, 1100, 3, 40, 2)]:
        p1 = torch.rand(N, P1, D, generator=gen).to(DEV)
        p2 = torch.rand(N, P2, D, generator=gen).to(DEV)
        l1 = torch.randint(0, P1 + 1, (N,), generator=gen).to(DEV)
        l2 = torch.randint(0, P2 + 1, (N,), generator=gen).to(DEV)
        ri, rd = _C.knn_points_idx(p1, p2, l1, l2, norm, K, -1)
        ks = _C.KnnSliced(p1, p2, l1, l2, norm, K)
        ks.idx.fill_(-7)
        ks.dists.fill_(-7.0)
        ks.prepare()
        cuts = [0, 1, 1, N // 2, N]  # includes an empty range
        for a, b in zip(cuts[:-1], cuts[1:]):
            ks.search(a, b)
        torch.cuda.synchronize()
        assert torch.equal(ks.idx, ri) and torch.equal(ks.dists, rd), (N, P1, P2, D, K, norm)
    # self-KNN through the same tensor (the pre-pass sorts the cloud once)
    p = torch.rand(4, 3000, 3, generator=gen).to(DEV)
    L = torch.tensor([3000, 1, 0, 2999], device=DEV)
    ri, rd = _C.knn_points_idx(p, p, L, L, 2, 16, -1)
    ks = _C.KnnSliced(p, p, L, L, 2, 16)
    ks.prepare()
    ks.search(2, 4)
    ks.search(0, 2)
    torch.cuda.synchronize()
    assert torch.equal(ks.idx, ri) and torch.equal(ks.dists, rd)
    with pytest.raises(RuntimeError):
        ks.search(3, 9)


def test_pair_search_matches_two_calls():
    """pops_knn_points_idx_pair = knn_points_idx(p1, p2) and knn_points_idx(p2, p1), bit for bit, on
    the shared-pre-pass path (both clouds >= 1024 points, D=3, L2) with ragged, tiny and empty
    clouds, duplicated points and offset clouds, and on the shapes where it is two plain calls."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(33)
    cases = [(6, 3000, 2200, 3, 1, 2), (4, 1024, 5000, 3, 8, 2), (3, 1500, 1500, 3, 32, 2), (3, 700, 2000, 3, 1, 2),
             (2, 1200, 1300, 3, 1, 1), (2, 300, 400, 5, 4, 2), (2, 1100, 9000, 3, 1, 2)]
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    # both pre-passes: the single-launch one (a thread-block cluster sorts the pair in distributed shared
    # memory) and the device-wide sort
    for fused, (N, P1, P2, D, K, norm) in [(f, c) for f in (1, 0) for c in cases]:
        lib.pops_set_option(b"knn_fused_prepass", fused)
        p1 = torch.rand(N, P1, D, generator=gen).to(DEV)
        p2 = (torch.rand(N, P2, D, generator=gen) * 0.7 + 0.4).to(DEV)  # partly overlapping boxes
        p2[:, : P2 // 8] = p2[:, P2 // 8 : 2 * (P2 // 8)]  # duplicated points
        l1 = torch.randint(1, P1 + 1, (N,), generator=gen).to(DEV)
        l2 = torch.randint(1, P2 + 1, (N,), generator=gen).to(DEV)
        l1[0], l2[0] = P1, P2
        if N > 2:
            l1[1], l2[1] = 3, P2   # tiny cloud against a full one
            l1[2], l2[2] = P1, 0   # empty cloud
        i12, d12, i21, d21 = _C.knn_points_idx_pair(p1, p2, l1, l2, norm, K)
        ri, rd = _C.knn_points_idx(p1, p2, l1, l2, norm, K, -1)
        si, sd = _C.knn_points_idx(p2, p1, l2, l1, norm, K, -1)
        torch.cuda.synchronize()
        assert torch.equal(i12, ri) and torch.equal(d12, rd), (N, P1, P2, D, K, norm)
        assert torch.equal(i21, si) and torch.equal(d21, sd), (N, P1, P2, D, K, norm)
    lib.pops_set_option(b"knn_fused_prepass", 1)


def _lib_opt(name, value):
    from pytorch3d_pointops_b200 import _lib

    _lib.load().pops_set_option(name, value)


@pytest.fixture
def force_ordered():
    """Route every D=3 L2 K<=32 call through the curve-ordered, box-pruned search, whatever P2."""
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    lib.pops_set_option(b"knn_order", 1)
    yield
    lib.pops_set_option(b"knn_order", -1)


def test_pruned_search_adversarial(oracle, golden, force_ordered):
    """The pruned path on the inputs that stress it: exact ties on integer grids (bounds equal to
    the K-th distance), duplicated points (degenerate boxes), offsets / scales (filter margins),
    collinear clouds (flat boxes), K > lengths2, empty clouds, tiny clouds, p1 != p2."""
    _C, _, _ = _ops()
    g = golden("knn_cases")
    p = g.t("grid.p", DEV)
    L = torch.tensor([p.shape[1]], device=DEV)
    for K in (1, 3, 7, 16, 17, 32):
        idx, dists = _C.knn_points_idx(p, p, L, L, 2, K, -1)
        assert torch.equal(idx.cpu(), g.t(f"grid.K{K}.idx")), K
        assert torch.equal(dists.cpu(), g.t(f"grid.K{K}.dists")), K
    gen = torch.Generator().manual_seed(78)
    base = torch.rand(2, 2500, 3, generator=gen)
    cases = {
        "dups": base[:, torch.randint(0, 40, (2500,), generator=gen)],
        "same": torch.full((2, 2500, 3), 0.5),
        "offset1e3": base + 1000.0,
        "offset1e5": base * 0.01 + 1e5,
        "tiny": base * 1e-20,
        "huge": base * 1e15,
        "mixed": torch.cat([base[:, :1200] * 1e-3, base[:, 1200:] * 50 + 7], 1),
        "line": torch.stack([base[..., 0], base[..., 0] * 0, base[..., 0] * 0], -1),
        "clusters": (base * 0.01 + torch.randint(0, 3, (2, 2500, 1), generator=gen).float()),
    }
    for name, pts in cases.items():
        pts = pts.contiguous()
        Lc = torch.tensor([2500, 1301])
        for K in (1, 16, 32):
            oi, od = oracle.knn_points_idx(pts, pts, Lc, Lc, 2, K, threads=8)
            for fused in (1, 0):
                _lib_opt(b"knn_fused_prepass", fused)
                gi, gd = _C.knn_points_idx(pts.to(DEV), pts.to(DEV), Lc.to(DEV), Lc.to(DEV), 2, K, -1)
                assert torch.equal(gi.cpu(), oi), (name, K, fused)
                assert torch.equal(gd.cpu(), od), (name, K, fused)
            _lib_opt(b"knn_fused_prepass", 1)
    # p1 != p2, ragged both sides, K > lengths2, empty and one-point clouds
    p1 = torch.randn(4, 700, 3, generator=gen)
    p2 = torch.randn(4, 1500, 3, generator=gen) * 0.5 + 0.3
    l1 = torch.tensor([700, 0, 13, 699])
    l2 = torch.tensor([1500, 900, 5, 0])
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    for K in (1, 8, 16, 20):
        oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, K)
        for fused in (1, 0):  # single-launch pre-pass / device-wide sort
            lib.pops_set_option(b"knn_fused_prepass", fused)
            gi, gd = _C.knn_points_idx(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), 2, K, -1)
            assert torch.equal(gi.cpu(), oi), (K, fused)
            assert torch.equal(gd.cpu(), od), (K, fused)
    lib.pops_set_option(b"knn_fused_prepass", 1)
    one = torch.rand(1, 1, 3, generator=gen)
    gi, gd = _C.knn_points_idx(one.to(DEV), one.to(DEV), torch.tensor([1], device=DEV), torch.tensor([1], device=DEV), 2, 4, -1)
    oi, od = oracle.knn_points_idx(one, one, torch.tensor([1]), torch.tensor([1]), 2, 4)
    assert torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od)


def test_pruned_search_both_thread_shapes(oracle):
    """Q = 2 and Q = 4 queries per thread, wide candidate ids (clouds beyond 262144 points use u32)."""
    from pytorch3d_pointops_b200 import _lib

    _C, _, _ = _ops()
    lib = _lib.load()
    gen = torch.Generator().manual_seed(5)
    p = torch.rand(2, 5000, 3, generator=gen)
    L = torch.tensor([5000, 3333])
    oi, od = oracle.knn_points_idx(p, p, L, L, 2, 16, threads=8)
    try:
        for q in (1, 2, 4):
            lib.pops_set_option(b"knn_q", q)
            gi, gd = _C.knn_points_idx(p.to(DEV), p.to(DEV), L.to(DEV), L.to(DEV), 2, 16, -1)
            assert torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od), q
    finally:
        lib.pops_set_option(b"knn_q", 0)
    big = torch.rand(1, 300000, 3, generator=gen)
    q1 = big[:, :256].contiguous()
    Lb, Lq = torch.tensor([300000]), torch.tensor([256])
    oi, od = oracle.knn_points_idx(q1, big, Lq, Lb, 2, 8, threads=8)
    gi, gd = _C.knn_points_idx(q1.to(DEV), big.to(DEV), Lq.to(DEV), Lb.to(DEV), 2, 8, -1)
    assert torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od)


@pytest.mark.gpu
@pytest.mark.parametrize("items", [0, 2, 4, 8])
def test_cluster_prepass_every_slice_size(oracle, items):
    """order_cluster_kernel (knn_order.cu): 1, 2, 4 or 8 CTAs per tensor, 2 / 4 / 8 keys per thread, 32- and
    64-bit sort keys, one or two tensors, ragged / empty / tiny clouds, duplicated points (equal codes
    straddling CTA borders), a flat cloud (one axis without extent).  Every configuration must give the
    oracle's rows and, on the larger shapes, the rows of the device-wide-sort pre-pass."""
    from pytorch3d_pointops_b200 import _lib

    _C, _, _ = _ops()
    lib = _lib.load()
    gen = torch.Generator().manual_seed(900 + items)
    try:
        lib.pops_set_option(b"knn_cluster_items", items)
        lib.pops_set_option(b"knn_order", 1)
        # small: against the oracle.  (N, P1, P2, K, self)
        for N, P1, P2, K, selfk in [(3, 700, 700, 8, True), (4, 3000, 2100, 4, False), (2, 5000, 5000, 16, True),
                                    (2, 100, 9000, 1, False), (1, 64, 64, 3, True)]:
            p2 = torch.rand(N, P2, 3, generator=gen)
            p2[:, : P2 // 4] = p2[:, P2 // 4 : 2 * (P2 // 4)]  # duplicates: long runs of equal codes
            p2[-1, :, 2] = 0.25                                 # a flat cloud
            l2 = torch.randint(1, P2 + 1, (N,), generator=gen)
            l2[0] = P2
            if selfk:
                p1, l1 = p2, l2
            else:
                p1 = torch.rand(N, P1, 3, generator=gen) * 1.3 - 0.1  # queries outside the blocks' box
                l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
                if N > 2:
                    l2[1] = 0  # nothing to search
            oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, K, threads=8)
            d1, g1 = p1.to(DEV), l1.to(DEV)
            d2, g2 = (d1, g1) if selfk else (p2.to(DEV), l2.to(DEV))  # same tensors: the self-search pre-pass
            gi, gd = _C.knn_points_idx(d1, d2, g1, g2, 2, K, -1)
            assert torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od), (items, N, P1, P2, K, selfk)
            if not selfk:
                i12, d12, i21, d21 = _C.knn_points_idx_pair(d1, d2, g1, g2, 2, K)
                ri, rd = oracle.knn_points_idx(p2, p1, l2, l1, 2, K, threads=8)
                assert torch.equal(i12.cpu(), oi) and torch.equal(d12.cpu(), od), (items, "pair 12")
                assert torch.equal(i21.cpu(), ri) and torch.equal(d21.cpu(), rd), (items, "pair 21")
        # large: against the device-wide sort (64-bit keys beyond 16384 positions per tensor)
        for N, P1, P2, K, selfk in [(2, 40000, 40000, 8, True), (2, 20000, 30000, 4, False), (3, 16384, 16384, 16, True)]:
            p2 = torch.rand(N, P2, 3, generator=gen).to(DEV)
            l2 = torch.randint(P2 // 2, P2 + 1, (N,), generator=gen).to(DEV)
            p1 = p2 if selfk else torch.rand(N, P1, 3, generator=gen).to(DEV)
            l1 = l2 if selfk else torch.randint(P1 // 2, P1 + 1, (N,), generator=gen).to(DEV)
            got = _C.knn_points_idx(p1, p2, l1, l2, 2, K, -1)
            lib.pops_set_option(b"knn_fused_prepass", 0)
            want = _C.knn_points_idx(p1, p2, l1, l2, 2, K, -1)
            lib.pops_set_option(b"knn_fused_prepass", 1)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), (items, N, P1, P2, K, selfk)
    finally:
        lib.pops_set_option(b"knn_cluster_items", 0)
        lib.pops_set_option(b"knn_order", -1)
        lib.pops_set_option(b"knn_fused_prepass", 1)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 4, 8, 16, 32])
def test_run_boxes_edge_lengths_and_clusters(oracle, K):
    """Blocks whose 16-point runs are partly or wholly padding (lengths just around run and block boundaries),
    points in tight far-apart clusters with exact duplicates (run boxes of zero extent, queries whose bound
    reaches exactly one cluster), every K bucket of the pruned search: idx and dists equal the oracle's."""
    _C, _, _ = _ops()
    gen = torch.Generator().manual_seed(1000 + K)
    lens = [64, 65, 79, 80, 81, 96, 111, 112, 113, 127, 128, 129, 191, 193, 1000, 1023, 1025, 2047]
    N, P = len(lens), max(lens)
    centres = torch.tensor([[0.0, 0.0, 0.0], [10.0, 0.0, 0.0], [0.0, -7.0, 3.0], [5.0, 5.0, 5.0]])
    which = torch.randint(0, 4, (N, P), generator=gen)
    p2 = centres[which] + 0.01 * torch.randn(N, P, 3, generator=gen)
    p2[:, 1::7] = p2[:, 0::7][:, : p2[:, 1::7].shape[1]]  # exact duplicates
    p1 = centres[torch.randint(0, 4, (N, 200), generator=gen)] + 0.5 * torch.randn(N, 200, 3, generator=gen)
    l2 = torch.tensor(lens)
    l1 = torch.randint(1, 201, (N,), generator=gen)
    oi, od = oracle.knn_points_idx(p1, p2, l1, l2, 2, K)
    gi, gd = _C.knn_points_idx(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), 2, K, -1)
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)
    # self-search of the same clouds (queries ARE the points: zero distances, ties among the duplicates)
    oi, od = oracle.knn_points_idx(p2, p2, l2, l2, 2, K)
    gi, gd = _C.knn_points_idx(p2.to(DEV), p2.to(DEV), l2.to(DEV), l2.to(DEV), 2, K, -1)
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)
