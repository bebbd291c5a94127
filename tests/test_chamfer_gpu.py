"""GPU parity: chamfer_distance (all reduction modes, features, weights, norms) against the
reference's golden outputs and gradients.  Losses/grads within 1e-5 relative."""
import pytest
import torch

from test_oracle import chamfer_variants, run_chamfer_variant

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_chamfer_all_variants_vs_golden(golden):
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    g = golden("chamfer_cases")
    for vi, v in enumerate(chamfer_variants(golden)):
        flat, grads = run_chamfer_variant(chamfer_distance, g, v, DEV)
        assert len(flat) == int(g.a(f"v{vi}.nout")), v
        n_loss = 2 if (v["pr"] is None and not v["sd"]) else 1
        for i, t in enumerate(flat):
            # distance terms: 1e-5 relative.  cosine terms are 1 - cos: one ulp of cos (6e-8) is an
            # ABSOLUTE error of the result, so they get atol 1e-6 on O(1) values.
            atol = 1e-7 if i < n_loss else 1e-6
            assert torch.allclose(t.detach().cpu(), g.t(f"v{vi}.out{i}"), rtol=1e-5, atol=atol), (v, i)
        for n, gr in grads.items():
            want = g.t(f"v{vi}.g_{n}")
            if want.numel() == 0:
                assert gr.numel() == 0 or not gr.any(), (v, n)
            else:
                # float atomics reorder the scatter-add: 1e-5 relative to the gradient scale
                assert torch.allclose(gr.cpu(), want, rtol=1e-5, atol=1e-5 * float(want.abs().max())), (v, n)


def test_chamfer_pointclouds_input(golden):
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
    from pytorch3d_pointops_b200.structures import Pointclouds

    g = golden("chamfer_cases")
    x, y, xl, yl = g.t("x", DEV), g.t("y", DEV), g.t("xl").tolist(), g.t("yl").tolist()
    xn, yn = g.t("xn", DEV), g.t("yn", DEV)
    pcx = Pointclouds([x[i, :l] for i, l in enumerate(xl)], features={"normals": [xn[i, :l] for i, l in enumerate(xl)]})
    pcy = Pointclouds([y[i, :l] for i, l in enumerate(yl)], features={"normals": [yn[i, :l] for i, l in enumerate(yl)]})
    loss, lf = chamfer_distance(pcx, pcy, feature_names=["normals"])
    assert torch.allclose(loss.cpu(), g.t("pc.loss"), rtol=1e-5)
    assert torch.allclose(lf["normals"].cpu(), g.t("pc.normals"), rtol=1e-5)
    # features present but feature_names None -> second element None (chamfer.py:107-112)
    loss2, lf2 = chamfer_distance(pcx, pcy)
    assert lf2 is None and torch.allclose(loss2, loss)
    with pytest.raises(ValueError, match="missing in x_features"):
        chamfer_distance(Pointclouds([x[0]]), Pointclouds([y[0]]), feature_names=["normals"])


def test_chamfer_config2_vs_oracle(oracle):
    """BASELINE.json configs[1] at reduced batch (B=4, P<=2048 ragged, normals+colors):
    loss, feature losses and all gradients against the oracle."""
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    gen = torch.Generator().manual_seed(1)
    N, P = 4, 2048
    x, y = torch.rand(N, P, 3, generator=gen), torch.rand(N, P, 3, generator=gen)
    xl = torch.randint(P // 2, P + 1, (N,), generator=gen)
    yl = torch.randint(P // 2, P + 1, (N,), generator=gen)
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


def test_chamfer_one_node_matches_per_direction_path():
    """The two-sided call as ONE autograd node (pair pre-pass, shared output tensor, accumulate-mode
    backward) against the per-direction nodes (forced by weights = 1): losses and all gradients,
    every batch / point reduction it serves, both norms, with and without features, ragged clouds
    including a 3-point one, and an upstream gradient that differs per output."""
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    gen = torch.Generator().manual_seed(5)
    N, P1, P2 = 5, 1500, 2100
    x, y = torch.rand(N, P1, 3, generator=gen), torch.rand(N, P2, 3, generator=gen) * 0.8 + 0.3
    xl = torch.randint(1024, P1 + 1, (N,), generator=gen)
    yl = torch.randint(1024, P2 + 1, (N,), generator=gen)
    xl[1], yl[2] = 3, 1
    xn, yn = torch.randn(N, P1, 3, generator=gen), torch.randn(N, P2, 3, generator=gen)
    xc, yc = torch.rand(N, P1, 4, generator=gen), torch.rand(N, P2, 4, generator=gen)
    ones = torch.ones(N, device=DEV)

    def run(weights, br, pr, norm, feats):
        ts = [t.to(DEV).clone().requires_grad_(True) for t in (x, y, xn, yn, xc, yc)]
        kw = dict(x_lengths=xl.to(DEV), y_lengths=yl.to(DEV), weights=weights, batch_reduction=br,
                  point_reduction=pr, norm=norm)
        if feats:
            kw.update(x_features={"normals": ts[2], "colors": ts[4]}, y_features={"normals": ts[3], "colors": ts[5]},
                      feature_names=["normals", "colors"])
        loss, lf = chamfer_distance(ts[0], ts[1], **kw)
        outs = [loss] + ([lf["normals"], lf["colors"]] if feats else [])
        up = [torch.linspace(0.5, 1.5, o.numel(), device=DEV).view(o.shape) * (i + 1) for i, o in enumerate(outs)]
        torch.autograd.backward(outs, up)
        return outs, [t.grad for t in (ts if feats else ts[:2])]

    for br in (None, "sum", "mean"):
        for pr in ("sum", "mean"):
            for norm, feats in ((2, True), (1, False), (2, False)):
                a_out, a_g = run(None, br, pr, norm, feats)
                b_out, b_g = run(ones, br, pr, norm, feats)
                for a, b in zip(a_out, b_out):
                    assert a.shape == b.shape and torch.allclose(a, b, rtol=2e-6, atol=1e-7), (br, pr, norm, feats)
                for a, b in zip(a_g, b_g):
                    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(b.abs().max())), (br, pr, norm, feats)
