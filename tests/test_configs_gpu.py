"""BASELINE.json configs at their full per-cloud sizes (fewer clouds where the CPU oracle would take
minutes): exact oracle agreement on the sampled part plus size-independent properties on the rest."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_config3_fps_full_cloud(oracle):
    """configs[2]: K=1024 from P=65536 (2 of the 64 clouds, one ragged): bit-exact index sequence."""
    from pytorch3d_pointops_b200.functions import sample_farthest_points

    gen = torch.Generator().manual_seed(2)
    pts = torch.rand(2, 65536, 3, generator=gen)
    L = torch.tensor([65536, 50001])
    _, oi = oracle.sample_farthest_points(pts, L, 1024)
    sp, gi = sample_farthest_points(pts.to(DEV), L.to(DEV), 1024)
    assert torch.equal(gi.cpu(), oi)
    assert gi[:, 0].eq(0).all() and (gi < L.to(DEV)[:, None]).all()
    assert all(len(set(row.tolist())) == 1024 for row in gi.cpu())  # no point is taken twice
    assert torch.equal(sp.cpu(), pts[torch.arange(2)[:, None], oi])


def test_config4_ball_query_sa_layer(oracle):
    """configs[3]: K=32, r=0.1, P=16384 (4 of the 128 clouds) + the gathered neighbourhoods."""
    from pytorch3d_pointops_b200.functions import ball_query

    gen = torch.Generator().manual_seed(3)
    N, P, K, r = 4, 16384, 32, 0.1
    p = torch.rand(N, P, 3, generator=gen)
    L = torch.tensor([P, P, 12000, 9000])
    pd, Ld = p.to(DEV), L.to(DEV)
    res = ball_query(pd, pd, Ld, Ld, K=K, radius=r, return_nn=True)
    idx, dists, nn = res.idx, res.dists, res.knn
    valid_q = torch.arange(P, device=DEV)[None] < Ld[:, None]
    filled = idx >= 0
    # padding: -1 / 0 / zero rows, and nothing for queries beyond lengths1
    assert not filled[~valid_q].any()
    assert not dists[~filled].any() and not nn[~filled].any()
    # hits: strictly inside the ball, ascending indices, inside lengths2, self included first... by index
    assert (dists[filled] < r * r).all()
    asc = (idx[..., 1:] > idx[..., :-1]) | ~filled[..., 1:]
    assert asc.all()
    assert (idx < Ld[:, None, None]).all()
    rows = torch.arange(N, device=DEV)[:, None, None]
    assert torch.equal(nn[filled], pd[rows.expand_as(idx)[filled], idx[filled]])
    # a query with fewer than K hits must have scanned everything: count equals the true count
    short = filled.sum(-1) < K
    n0 = 1
    q = short[n0].nonzero()[:16, 0]
    d2 = ((pd[n0, q, None, :] - pd[n0, None, : int(L[n0]), :]) ** 2)
    true_cnt = ((d2[..., 0] + d2[..., 1]) + d2[..., 2]).lt(r * r).sum(-1)
    assert torch.equal(filled[n0, q].sum(-1), true_cnt)
    # exact oracle agreement on slices of queries of a full and a ragged cloud
    for n, q0 in ((0, 5000), (3, 8990)):
        oi, od = oracle.ball_query_idx(p[n:n + 1], p[n:n + 1], L[n:n + 1], L[n:n + 1], K, r, q0=q0, q1=q0 + 64, threads=8)
        assert torch.equal(idx[n, q0:q0 + 64].cpu(), oi[0, q0:q0 + 64])
        assert torch.equal(dists[n, q0:q0 + 64].cpu(), od[0, q0:q0 + 64])


def test_config2_chamfer_full_cloud_size(oracle):
    """configs[1] at P<=8192 (2 of the 32 cloud pairs): loss, feature losses and gradients."""
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    gen = torch.Generator().manual_seed(1)
    N, P = 2, 8192
    x, y = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)
    xl, yl = torch.tensor([8192, 4100]), torch.tensor([5003, 8192])
    xn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=gen), dim=-1)
    yn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=gen), dim=-1)
    xc, yc = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)

    def run(fn, dev):
        ts = [t.to(dev).clone().requires_grad_(True) for t in (x, y, xn, yn, xc, yc)]
        loss, lf = fn(ts[0], ts[1], x_lengths=xl.to(dev), y_lengths=yl.to(dev),
                      x_features={"normals": ts[2], "colors": ts[4]},
                      y_features={"normals": ts[3], "colors": ts[5]},
                      feature_names=["normals", "colors"])
        (loss + lf["normals"] + lf["colors"]).backward()
        return [loss, lf["normals"], lf["colors"]], [t.grad for t in ts]

    o_out, o_grads = run(oracle.chamfer_distance, "cpu")
    g_out, g_grads = run(chamfer_distance, DEV)
    for a, b in zip(g_out, o_out):
        assert torch.allclose(a.detach().cpu(), b.detach(), rtol=1e-5, atol=1e-8)
    for a, b in zip(g_grads, o_grads):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))
