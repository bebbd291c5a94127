"""GPU parity: ball_query, sample_farthest_points, packed<->padded, Pointclouds plumbing."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------ ball query
@pytest.mark.parametrize("name", ["a", "b", "c", "none"])
def test_ball_query_golden(golden, name):
    from pytorch3d_pointops_b200.functions import ball_query

    g = golden("ball_query_cases")
    p1 = g.t(f"{name}.p1", DEV).requires_grad_(True)
    p2 = g.t(f"{name}.p2", DEV).requires_grad_(True)
    K, radius = int(g.a(f"{name}.K")), float(g.a(f"{name}.radius"))
    res = ball_query(p1, p2, g.t(f"{name}.l1", DEV), g.t(f"{name}.l2", DEV), K=K, radius=radius)
    assert torch.equal(res.idx.cpu(), g.t(f"{name}.idx"))
    assert torch.equal(res.dists.detach().cpu(), g.t(f"{name}.dists"))
    assert torch.equal(res.knn.detach().cpu(), g.t(f"{name}.knn"))
    ((res.dists * g.t(f"{name}.gd", DEV)).sum() + (res.knn * g.t(f"{name}.gn", DEV)).sum()).backward()
    for got, want in ((p1.grad.cpu(), g.t(f"{name}.grad_p1")), (p2.grad.cpu(), g.t(f"{name}.grad_p2"))):
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * max(1.0, float(want.abs().max())))


def test_ball_query_grid_ties(golden):
    """Grid points exactly on the radius: strict '<' on the float32 product radius*radius."""
    from pytorch3d_pointops_b200.functions import ball_query

    g = golden("ball_query_cases")
    p = g.t("grid.p", DEV)
    res = ball_query(p, p, K=30, radius=0.25)
    assert torch.equal(res.idx.cpu(), g.t("grid.idx"))
    assert torch.equal(res.dists.cpu(), g.t("grid.dists"))


@pytest.mark.parametrize("N,P1,P2,D,K,radius", [
    (2, 3000, 5000, 3, 32, 0.1), (3, 513, 2049, 3, 8, 0.2), (1, 1000, 1000, 3, 500, 0.3),
    (2, 200, 300, 2, 16, 0.15), (1, 100, 400, 6, 10, 0.8), (2, 10, 5000, 3, 1, 0.05),
])
def test_ball_query_oracle_sweep(oracle, N, P1, P2, D, K, radius):
    from pytorch3d_pointops_b200 import _C

    gen = torch.Generator().manual_seed(P1 + P2 + K)
    p1 = torch.rand(N, P1, D, generator=gen)
    p2 = torch.rand(N, P2, D, generator=gen)
    l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
    l2 = torch.randint(0, P2 + 1, (N,), generator=gen)
    l1[0], l2[0] = P1, P2
    oi, od = oracle.ball_query_idx(p1, p2, l1, l2, K, radius)
    gi, gd = _C.ball_query(p1.to(DEV), p2.to(DEV), l1.to(DEV), l2.to(DEV), K, radius)
    assert torch.equal(gi.cpu(), oi)
    assert torch.equal(gd.cpu(), od)


def test_ball_query_offsets(oracle):
    from pytorch3d_pointops_b200 import _C

    gen = torch.Generator().manual_seed(9)
    for off, scale in ((1000.0, 1.0), (0.0, 1e-10), (-3e4, 10.0)):
        p = (torch.rand(1, 2000, 3, generator=gen) * scale + off).contiguous()
        r = 0.1 * scale
        oi, od = oracle.ball_query_idx(p, p, None, None, 16, r)
        L = torch.tensor([2000], device=DEV)
        gi, gd = _C.ball_query(p.to(DEV), p.to(DEV), L, L, 16, r)
        assert torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od)


@pytest.fixture
def bq_spatial_forced():
    """Every finite cloud of an eligible shape takes the Hilbert-ordered search (bq_prune.cu)."""
    from pytorch3d_pointops_b200 import _lib

    lib = _lib.load()
    lib.pops_set_option(b"bq_spatial", 1)
    yield lib
    lib.pops_set_option(b"bq_spatial", -1)


def _bq_both(oracle, p1, p2, l1, l2, K, r, self_search=False):
    from pytorch3d_pointops_b200 import _C

    oi, od = oracle.ball_query_idx(p1, p2, l1, l2, K, r)
    d1, g1 = p1.to(DEV), l1.to(DEV)
    d2, g2 = (d1, g1) if self_search else (p2.to(DEV), l2.to(DEV))
    gi, gd = _C.ball_query(d1, d2, g1, g2, K, r)
    assert torch.equal(gi.cpu(), oi), (K, r)
    assert torch.equal(gd.cpu(), od), (K, r)


def test_ball_query_spatial_vs_oracle(oracle, bq_spatial_forced):
    """bq_prune_kernel on what stresses it: sparse and crowded balls (hit columns cut back to K many
    times), K = 1 .. 64, ragged / empty / tiny clouds, p1 != p2 with queries outside the cloud's box,
    duplicated points, a flat cloud, clouds far from the origin (filter margin), and the per-cloud
    dispatch leaving non-finite clouds to the exact scan."""
    gen = torch.Generator().manual_seed(77)
    for N, P1, P2, K, r, selfk in [(3, 2500, 2500, 32, 0.1, True), (2, 1024, 6000, 64, 0.15, False),
                                   (2, 3000, 2048, 1, 0.05, False), (2, 4000, 4000, 16, 0.4, True),
                                   (1, 1100, 9000, 5, 0.03, False), (2, 2100, 2100, 64, 1.5, True)]:
        p2 = torch.rand(N, P2, 3, generator=gen)
        p2[:, : P2 // 8] = p2[:, P2 // 8 : 2 * (P2 // 8)]  # duplicated points: equal distances, equal codes
        p2[-1, :, 1] = 0.5                                  # a flat cloud
        l2 = torch.randint(P2 // 2, P2 + 1, (N,), generator=gen)
        if selfk:
            p1, l1 = p2, l2
        else:
            p1 = torch.rand(N, P1, 3, generator=gen) * 1.4 - 0.2
            l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
            l1[0] = P1
            if N > 1:
                l2[1] = 0  # nothing to find
        _bq_both(oracle, p1, p2, l1, l2, K, r, selfk)
    # offsets and scales: the expanded-form filter must keep every true hit
    for off, scale in ((1000.0, 1.0), (0.0, 1e-6), (-3e4, 10.0)):
        p = (torch.rand(2, 3000, 3, generator=gen) * scale + off).contiguous()
        L = torch.tensor([3000, 2222])
        _bq_both(oracle, p, p, L, L, 16, 0.12 * scale, True)
    # one cloud non-finite, one huge: both answered by the exact scan, the others by the ordered search
    p = torch.rand(4, 2600, 3, generator=gen)
    p[1, 7, 2] = float("nan")
    p[2, :, :] *= 3e19
    L = torch.tensor([2600, 2600, 2600, 2000])
    _bq_both(oracle, p, p, L, L, 8, 0.1, True)


def test_ball_query_spatial_ties_on_the_radius(oracle, bq_spatial_forced):
    """Integer-grid cloud with the radius on exact inter-point distances: strict '<' on fl(r * r), with
    block-box lower bounds equal to r^2 (a block at exactly the radius holds no hit), many equal distances."""
    ax = torch.arange(14, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).contiguous()  # 2744 points
    perm = torch.randperm(grid.shape[1], generator=torch.Generator().manual_seed(3))
    grid = grid[:, perm].contiguous()
    L = torch.tensor([grid.shape[1]])
    for r in (1.0, 2.0, 3.0, 2.2360679):
        for K in (7, 33, 64):
            _bq_both(oracle, grid, grid, L, L, K, r, True)


def test_ball_query_auto_dispatch_mixed_batch(oracle):
    """Default dispatch on a batch whose clouds differ: a sparse cloud (ordered search), a crowded one and a
    short one (index scan) in one call; each row equals the oracle's whichever kernel produced it."""
    gen = torch.Generator().manual_seed(5)
    p = torch.rand(3, 6000, 3, generator=gen)
    p[1] = p[1] * 0.05 + 0.4  # 8000x denser: every ball is crowded
    L = torch.tensor([6000, 6000, 1500])
    _bq_both(oracle, p, p, L, L, 32, 0.04, True)


@pytest.mark.parametrize("P2", [70000])
def test_ball_query_spatial_wide_indices(oracle, bq_spatial_forced, P2):
    """Clouds beyond 65536 points keep 32-bit hit indices."""
    gen = torch.Generator().manual_seed(11)
    p2 = torch.rand(1, P2, 3, generator=gen)
    p1 = p2[:, -1500:].contiguous()
    _bq_both(oracle, p1, p2, torch.tensor([1500]), torch.tensor([P2]), 24, 0.03)


# ------------------------------------------------------------------ farthest point sampling
def test_fps_golden(golden):
    from pytorch3d_pointops_b200 import _C
    from pytorch3d_pointops_b200.functions import sample_farthest_points

    g = golden("fps_cases")
    sp, si = sample_farthest_points(g.t("ragged.points", DEV), g.t("ragged.lengths", DEV),
                                    g.t("ragged.K", DEV))
    assert torch.equal(si.cpu(), g.t("ragged.idx")) and torch.equal(sp.cpu(), g.t("ragged.sampled"))
    sp, si = sample_farthest_points(g.t("big.points", DEV), K=200)
    assert torch.equal(si.cpu(), g.t("big.idx")) and torch.equal(sp.cpu(), g.t("big.sampled"))
    _, si = sample_farthest_points(g.t("dup.points", DEV), K=4)
    assert si.cpu().tolist() == [[0, 0, 0, 0]]
    si = _C.sample_farthest_points(g.t("start.points", DEV), g.t("start.lengths", DEV),
                                   g.t("start.K", DEV), g.t("start.start", DEV))
    assert torch.equal(si.cpu(), g.t("start.idx"))
    _, si = sample_farthest_points(g.t("d6.points", DEV), K=33)  # generic-D kernel
    assert torch.equal(si.cpu(), g.t("d6.idx"))


@pytest.mark.parametrize("N,P,K", [(1, 100, 100), (2, 1500, 64), (3, 5000, 300), (1, 20000, 128),
                                   (2, 40000, 50), (64, 1024, 32), (1, 70000, 40), (1, 131072, 16)])
def test_fps_oracle_sweep(oracle, N, P, K):
    """Covers every cluster size (1..16 CTAs per cloud) and points-per-thread variant."""
    from pytorch3d_pointops_b200.functions import sample_farthest_points

    gen = torch.Generator().manual_seed(P + K)
    pts = torch.rand(N, P, 3, generator=gen)
    L = torch.randint(max(1, P // 2), P + 1, (N,), generator=gen)
    L[0] = P
    _, oi = oracle.sample_farthest_points(pts, L, K)
    sp, gi = sample_farthest_points(pts.to(DEV), L.to(DEV), K)
    assert torch.equal(gi.cpu(), oi)
    want = pts[torch.arange(N)[:, None], oi.clamp(min=0)] * oi.ne(-1)[..., None]
    assert torch.equal(sp.cpu(), want)


def test_fps_ties_and_quirks(oracle):
    from pytorch3d_pointops_b200 import _C

    # integer grid: massive exact ties in the arg-max -> lowest index must win
    grid = torch.stack(torch.meshgrid(*[torch.arange(12.0)] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    L = torch.tensor([grid.shape[1]])
    K = torch.tensor([200])
    s = torch.tensor([77])
    want = oracle.sample_farthest_points_idx(grid, L, K, s)
    got = _C.sample_farthest_points(grid.to(DEV), L.to(DEV), K.to(DEV), s.to(DEV))
    assert torch.equal(got.cpu(), want)
    # K[n] = 0 still writes the start index (sample_farthest_points_cpu.cpp:53-54)
    pts = torch.rand(2, 10, 3)
    got = _C.sample_farthest_points(pts.to(DEV), torch.tensor([10, 10], device=DEV),
                                    torch.tensor([0, 3], device=DEV), torch.tensor([4, 2], device=DEV))
    want = oracle.sample_farthest_points_idx(pts, torch.tensor([10, 10]), torch.tensor([0, 3]),
                                             torch.tensor([4, 2]))
    assert torch.equal(got.cpu(), want) and got[0].tolist() == [4, -1, -1]


def test_fps_random_start_matches_reference_rng():
    from pytorch3d_pointops_b200.functions import sample_farthest_points

    pts = torch.rand(3, 50, 3, device=DEV)
    L = torch.tensor([50, 20, 7], device=DEV)
    torch.manual_seed(123)
    _, a = sample_farthest_points(pts, L, K=5, random_start_point=True)
    torch.manual_seed(123)
    want = [int(torch.randint(high=h, size=(1,)).item()) for h in (50, 20, 7)]
    assert a[:, 0].tolist() == want


# ------------------------------------------------------------------ packed <-> padded
def test_packed_padded_golden(golden):
    from pytorch3d_pointops_b200.functions import packed_to_padded, padded_to_packed

    g = golden("packed_padded_cases")
    first, sizes = g.t("first", DEV), g.t("sizes")
    for D in (3, 1, 6):
        packed = g.t(f"D{D}.packed", DEV).requires_grad_(True)
        padded = packed_to_padded(packed, first, int(sizes.max()))
        assert torch.equal(padded.detach().cpu(), g.t(f"D{D}.padded"))
        (padded * g.t(f"D{D}.gpad", DEV)).sum().backward()
        assert torch.equal(packed.grad.cpu(), g.t(f"D{D}.grad_packed"))
        back = padded_to_packed(padded.detach(), first, int(sizes.sum()))
        assert torch.equal(back.cpu(), g.t(f"D{D}.packed"))
    # max_size_dim handling and empty clouds
    x = torch.randn(3, 4, 7, device=DEV)
    f = torch.tensor([0, 5, 5], device=DEV)
    out = padded_to_packed(x, f, 9, max_size_dim=2)
    want = torch.cat([x[0, :, :5].T, x[2, :, :4].T])
    assert torch.equal(out, want)


def test_packed_padded_oracle_ragged(oracle):
    from pytorch3d_pointops_b200 import _C

    gen = torch.Generator().manual_seed(4)
    sizes = torch.randint(0, 3000, (37,), generator=gen)
    sizes[5] = 0
    first = torch.zeros_like(sizes)
    first[1:] = sizes.cumsum(0)[:-1]
    F = int(sizes.sum())
    x = torch.randn(F, 5, generator=gen)
    want = oracle.packed_to_padded_C(x, first, int(sizes.max()))
    got = _C.packed_to_padded(x.to(DEV), first.to(DEV), int(sizes.max()))
    assert torch.equal(got.cpu(), want)
    back = _C.padded_to_packed(got, first.to(DEV), F)
    assert torch.equal(back.cpu(), x)
    assert torch.equal(back.cpu(), oracle.padded_to_packed_C(want, first, F))


# ------------------------------------------------------------------ Pointclouds on device
def test_pointclouds_cuda_plumbing(golden):
    from pytorch3d_pointops_b200.structures import Pointclouds

    g = golden("pointclouds_cases")
    sizes = g.t("sizes").tolist()
    pc = Pointclouds([g.t(f"pts{i}", DEV) for i in range(len(sizes))],
                     features={"normals": [g.t(f"nrm{i}", DEV) for i in range(len(sizes))],
                               "colors": [g.t(f"col{i}", DEV) for i in range(len(sizes))]})
    assert torch.equal(pc.points_padded().cpu(), g.t("points_padded"))
    assert torch.equal(pc.points_packed().cpu(), g.t("points_packed"))
    assert torch.equal(pc.features_padded()["normals"].cpu(), g.t("normals_padded"))
    assert torch.equal(pc.features_packed()["colors"].cpu(), g.t("colors_packed"))
    assert torch.equal(pc.padded_to_packed_idx().cpu(), g.t("padded_to_packed_idx"))
    assert torch.equal(pc.packed_to_cloud_idx().cpu(), g.t("packed_to_cloud_idx"))
    back = pc.cpu()
    assert torch.equal(back.points_packed(), g.t("points_packed"))


def test_sample_pdf_golden_and_oracle(golden, oracle):
    """sample_pdf: bit-exact vs the reference's golden vectors and the oracle; API errors."""
    from pytorch3d_pointops_b200 import _C
    from pytorch3d_pointops_b200.functions.sample_pdf import sample_pdf, sample_pdf_python

    g = golden("sample_pdf_cases")
    for name, n_samples in (("a", 33), ("b", 128), ("c", 9), ("d", 17)):
        bins, w = g.t(f"{name}.bins", DEV), g.t(f"{name}.weights", DEV)
        assert torch.equal(sample_pdf(bins, w, n_samples, det=True).cpu(), g.t(f"{name}.det"))
        out = g.t(f"{name}.u", DEV).clone()
        v0 = out._version
        n_bins = w.shape[-1]
        _C.sample_pdf(bins.reshape(-1, n_bins + 1), w.reshape(-1, n_bins), out.view(-1, n_samples), 1e-5)
        assert out._version > v0  # in place + version bump, like the reference
        assert torch.equal(out.cpu(), g.t(f"{name}.rand"))
    gen = torch.Generator().manual_seed(12)
    B, n_bins, n_samples = 4097, 64, 128  # NeRF-like: rays x bins (and a batch the reference cannot split safely)
    bins = torch.sort(torch.rand(B, n_bins + 1, generator=gen) * 6, dim=-1).values
    w = torch.rand(B, n_bins, generator=gen) * (torch.rand(B, n_bins, generator=gen) > 0.3)
    u = torch.rand(B, n_samples, generator=gen)
    want = oracle.sample_pdf(bins, w, n_samples, u=u)
    got = u.to(DEV).clone()
    _C.sample_pdf(bins.to(DEV), w.to(DEV), got, 1e-5)
    assert torch.equal(got.cpu(), want)
    # random sampling stays inside the support; the searchsorted variant agrees up to rounding
    s = sample_pdf(bins.to(DEV), w.to(DEV), 32)
    assert s.shape == (B, 32) and (s >= bins.min()).all() and (s <= bins.max()).all()
    det = sample_pdf(bins.to(DEV), w.to(DEV) + 0.05, 16, det=True)
    assert torch.allclose(det, sample_pdf_python(bins.to(DEV), w.to(DEV) + 0.05, 16, det=True), atol=2e-3)
    with pytest.raises(ValueError, match="Negative weights"):
        sample_pdf(bins.to(DEV), -w.to(DEV) - 1.0, 4)
    with pytest.raises(ValueError, match="Inconsistent shapes"):
        sample_pdf(bins.to(DEV)[:, :-1], w.to(DEV), 4)
    with pytest.raises(NotImplementedError):
        sample_pdf(bins.to(DEV).requires_grad_(True), w.to(DEV), 4)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _C.sample_pdf(bins, w, u.clone(), 1e-5)


def test_point_covariances_fused_matches_torch_formula():
    """get_point_covariances: the fused kernel (no grad) against the reference's torch formula
    (functions/utils.py:111-153) evaluated on our knn_points(return_nn=True)."""
    from pytorch3d_pointops_b200.functions import get_point_covariances, knn_points

    gen = torch.Generator().manual_seed(6)
    for D, K in ((3, 16), (2, 5), (4, 9)):
        pts = torch.randn(3, 700, D, generator=gen).to(DEV)
        L = torch.tensor([700, 433, 6], device=DEV)  # the last cloud has fewer points than K=9/16
        cov, nn = get_point_covariances(pts, L, K)
        ref_nn = knn_points(pts, pts, lengths1=L, lengths2=L, K=K, return_nn=True).knn
        centred = ref_nn - ref_nn.mean(2, keepdim=True)
        ref_cov = (centred.unsqueeze(4) * centred.unsqueeze(3)).mean(2)
        assert torch.equal(nn, ref_nn)
        assert torch.allclose(cov, ref_cov, rtol=1e-5, atol=1e-6)
        # with gradients the torch path is taken and stays differentiable
        pg = pts.clone().requires_grad_(True)
        cov_g, _ = get_point_covariances(pg, L, K)
        cov_g.sum().backward()
        assert torch.allclose(cov_g.detach(), cov, rtol=1e-5, atol=1e-6) and pg.grad is not None
