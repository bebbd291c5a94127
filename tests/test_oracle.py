"""CPU: the oracle restatement (oracle/) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and, when present, against oracle/_ref itself."""
import ast

import pytest
import torch

KNN_CASES = ["d3k1", "d3k16", "d3k32_l1", "d3k40", "d2k5", "d5k7", "d33k4", "d128k16", "klen"]


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_forward_backward_vs_golden(golden, oracle, name):
    g = golden("knn_cases")
    p1, p2, l1, l2 = g.t(f"{name}.p1"), g.t(f"{name}.p2"), g.t(f"{name}.l1"), g.t(f"{name}.l2")
    K, norm = int(g.a(f"{name}.K")), int(g.a(f"{name}.norm"))
    idx, dists = oracle.knn_points_idx(p1, p2, l1, l2, norm, K)
    assert torch.equal(idx, g.t(f"{name}.idx"))
    assert torch.equal(dists, g.t(f"{name}.dists"))
    nn = oracle.knn_gather(p2, idx, l2)
    assert torch.equal(nn, g.t(f"{name}.knn"))
    # backward: d(dists*gd + knn*gn)
    p1r, p2r = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
    d, i, n_ = oracle.knn_points(p1r, p2r, l1, l2, norm=norm, K=K, return_nn=True)
    ((d * g.t(f"{name}.gd")).sum() + (n_ * g.t(f"{name}.gn")).sum()).backward()
    assert torch.allclose(p1r.grad, g.t(f"{name}.grad_p1"), rtol=1e-6, atol=1e-6)
    assert torch.allclose(p2r.grad, g.t(f"{name}.grad_p2"), rtol=1e-5, atol=1e-5)


def test_knn_readme_config(golden, oracle):
    g = golden("knn_readme")
    X, L = g.t("padded"), g.t("lengths")
    idx, dists = oracle.knn_points_idx(X, X, L, L, 2, 8)
    assert torch.equal(idx, g.t("idx"))
    assert torch.equal(dists, g.t("dists"))
    assert abs(float(dists.sum()) - 2895.012695) < 1e-2  # SURVEY.md section 4 golden
    assert idx[1, 799].tolist() == [799, 637, 38, 426, 705, 359, 524, 798]
    assert not idx[1, 800:].any() and not dists[1, 800:].any()


@pytest.mark.parametrize("K", [1, 3, 7, 16, 17, 32, 33])
def test_knn_ties_grid(golden, oracle, K):
    g = golden("knn_cases")
    p = g.t("grid.p")
    idx, dists = oracle.knn_points_idx(p, p, None, None, 2, K)
    assert torch.equal(idx, g.t(f"grid.K{K}.idx"))
    assert torch.equal(dists, g.t(f"grid.K{K}.dists"))


def test_knn_tie_vector(golden, oracle):
    g = golden("knn_cases")
    for K, want in ((3, [0, 1, 2]), (4, [0, 1, 2, 3]), (5, [0, 1, 2, 3, 4])):
        idx, _ = oracle.knn_points_idx(g.t("tie.p1"), g.t("tie.p2"), None, None, 2, K)
        assert idx[0, 0].tolist() == want
        assert torch.equal(idx, g.t(f"tie.K{K}.idx"))


@pytest.mark.parametrize("name", ["a", "b", "c", "none"])
def test_ball_query_vs_golden(golden, oracle, name):
    g = golden("ball_query_cases")
    p1, p2, l1, l2 = g.t(f"{name}.p1"), g.t(f"{name}.p2"), g.t(f"{name}.l1"), g.t(f"{name}.l2")
    K, radius = int(g.a(f"{name}.K")), float(g.a(f"{name}.radius"))
    dists, idx, nn = oracle.ball_query(p1, p2, l1, l2, K=K, radius=radius)
    assert torch.equal(idx, g.t(f"{name}.idx"))
    assert torch.equal(dists, g.t(f"{name}.dists"))
    assert torch.equal(nn, g.t(f"{name}.knn"))
    # backward through the dists (reference reuses knn backward with norm 2) and the gather
    g1, g2 = oracle.knn_points_backward(p1, p2, l1, l2, idx, 2, g.t(f"{name}.gd"))
    gn = g.t(f"{name}.gn")
    scatter = torch.zeros_like(p2)
    safe = idx.clamp(min=0)
    contrib = gn * idx.ne(-1)[..., None]
    scatter.scatter_add_(1, safe.reshape(p2.shape[0], -1, 1).expand(-1, -1, p2.shape[2]),
                         contrib.reshape(p2.shape[0], -1, p2.shape[2]))
    assert torch.allclose(g1, g.t(f"{name}.grad_p1"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(g2 + scatter, g.t(f"{name}.grad_p2"), rtol=1e-5, atol=1e-5)


def test_ball_query_grid(golden, oracle):
    g = golden("ball_query_cases")
    p = g.t("grid.p")
    dists, idx, _ = oracle.ball_query(p, p, K=30, radius=0.25)
    assert torch.equal(idx, g.t("grid.idx")) and torch.equal(dists, g.t("grid.dists"))


def test_fps_vs_golden(golden, oracle):
    g = golden("fps_cases")
    sp, si = oracle.sample_farthest_points(g.t("ragged.points"), g.t("ragged.lengths"),
                                           g.t("ragged.K"))
    assert torch.equal(si, g.t("ragged.idx")) and torch.equal(sp, g.t("ragged.sampled"))
    assert si[1].tolist()[:8] == [0, 1, 2, 6, 4, 5, 3, -1]  # SURVEY.md section 4 golden
    sp, si = oracle.sample_farthest_points(g.t("big.points"), K=200)
    assert torch.equal(si, g.t("big.idx")) and torch.equal(sp, g.t("big.sampled"))
    _, si = oracle.sample_farthest_points(g.t("dup.points"), K=4)
    assert torch.equal(si, g.t("dup.idx")) and si.tolist() == [[0, 0, 0, 0]]
    si = oracle.sample_farthest_points_idx(g.t("start.points"), g.t("start.lengths"),
                                           g.t("start.K"), g.t("start.start"))
    assert torch.equal(si, g.t("start.idx"))
    _, si = oracle.sample_farthest_points(g.t("d6.points"), K=33)
    assert torch.equal(si, g.t("d6.idx"))


def test_packed_padded_vs_golden(golden, oracle):
    g = golden("packed_padded_cases")
    first, sizes = g.t("first"), g.t("sizes")
    for D in (3, 1, 6):
        packed, padded = g.t(f"D{D}.packed"), g.t(f"D{D}.padded")
        flat = packed.reshape(packed.shape[0], -1)
        out = oracle.packed_to_padded_C(flat, first, int(sizes.max()))
        assert torch.equal(out.reshape(padded.shape), padded)
        back = oracle.padded_to_packed_C(out, first, flat.shape[0])
        assert torch.equal(back, flat)
        gback = oracle.padded_to_packed_C(g.t(f"D{D}.gpad").reshape(out.shape), first, flat.shape[0])
        assert torch.equal(gback.reshape(packed.shape), g.t(f"D{D}.grad_packed"))


def test_gathers_vs_golden(golden, oracle):
    g = golden("gather_cases")
    x = g.t("kg.x")
    assert torch.equal(oracle.knn_gather(x, g.t("kg.idx"), g.t("kg.lengths")), g.t("kg.out"))
    assert torch.equal(oracle.knn_gather(x, g.t("kg.idx")), g.t("kg.out_full"))
    assert torch.equal(oracle.masked_gather(x, g.t("mg.idx3")), g.t("mg.out3"))
    assert torch.equal(oracle.masked_gather(x, g.t("mg.idx2")), g.t("mg.out2"))


def _walk(o, out):
    if o is None:
        return
    if torch.is_tensor(o):
        out.append(o)
    elif isinstance(o, dict):
        for k in sorted(o):
            _walk(o[k], out)
    else:
        for e in o:
            _walk(e, out)


def chamfer_variants(golden):
    g = golden("chamfer_cases")
    return [ast.literal_eval(str(s)) for s in g.a("variants")]


def run_chamfer_variant(fn, g, v, device="cpu"):
    """Shared by the oracle test and the GPU parity test: returns (outputs, grads dict)."""
    names = ("x", "y", "xn", "yn", "xc", "yc")
    ts = {n: g.t(n, device).clone().requires_grad_(True) for n in names}
    kw = dict(batch_reduction=v["br"], point_reduction=v["pr"], norm=v["norm"],
              single_directional=v["sd"], abs_cosine=v["abs"])
    if v["ragged"]:
        kw.update(x_lengths=g.t("xl", device), y_lengths=g.t("yl", device))
    if v["w"]:
        kw["weights"] = g.t("w", device)
    if v["feats"]:
        kw.update(x_features={"normals": ts["xn"], "colors": ts["xc"]},
                  y_features={"normals": ts["yn"], "colors": ts["yc"]},
                  feature_names=["normals", "colors"])
    loss, lf = fn(ts["x"], ts["y"], **kw)
    flat = []
    _walk(loss, flat)
    _walk(lf, flat)
    total = sum((t * (i + 1)).sum() for i, t in enumerate(flat))
    total.backward()
    grads = {n: (ts[n].grad if ts[n].grad is not None else torch.zeros(0)) for n in names}
    return flat, grads


def test_chamfer_vs_golden(golden, oracle):
    g = golden("chamfer_cases")
    for vi, v in enumerate(chamfer_variants(golden)):
        flat, grads = run_chamfer_variant(oracle.chamfer_distance, g, v)
        assert len(flat) == int(g.a(f"v{vi}.nout")), v
        for i, t in enumerate(flat):
            assert torch.allclose(t, g.t(f"v{vi}.out{i}"), rtol=1e-6, atol=1e-7), (v, i)
        for n, gr in grads.items():
            want = g.t(f"v{vi}.g_{n}")
            if want.numel() == 0:
                assert gr.numel() == 0 or not gr.any(), (v, n)
            else:
                assert torch.allclose(gr, want, rtol=1e-5, atol=1e-6), (v, n)


def test_oracle_matches_compiled_reference(oracle, ref_module):
    """Seeded random inputs through oracle/_ref (the reference's own CPU sources) when present."""
    if ref_module is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt file)")
    R = ref_module
    gen = torch.Generator().manual_seed(5)
    for (N, P1, P2, D, K, norm) in [(2, 90, 70, 3, 8, 2), (2, 33, 65, 3, 16, 1), (1, 20, 40, 7, 5, 2),
                                    (2, 16, 4, 3, 6, 2), (1, 12, 30, 64, 3, 2)]:
        p1 = torch.randn(N, P1, D, generator=gen)
        p2 = torch.randn(N, P2, D, generator=gen)
        l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
        l2 = torch.randint(0, P2 + 1, (N,), generator=gen)
        ri, rd = R.knn_points_idx(p1, p2, l1, l2, norm, K, -1)
        oi, od = oracle.knn_points_idx(p1, p2, l1, l2, norm, K)
        assert torch.equal(ri, oi) and torch.equal(rd, od)
        gd = torch.randn(N, P1, K, generator=gen)
        rg = R.knn_points_backward(p1, p2, l1, l2, ri, norm, gd)
        og = oracle.knn_points_backward(p1, p2, l1, l2, oi, norm, gd)
        assert torch.equal(rg[0], og[0]) and torch.equal(rg[1], og[1])
        ri, rd = R.ball_query(p1, p2, l1, l2, K, 0.9)
        oi, od = oracle.ball_query_idx(p1, p2, l1, l2, K, 0.9)
        assert torch.equal(ri, oi) and torch.equal(rd, od)
    pts = torch.rand(3, 120, 3, generator=gen)
    L = torch.tensor([120, 40, 3])
    K = torch.tensor([30, 50, 2])
    s = torch.tensor([5, 0, 2])
    assert torch.equal(R.sample_farthest_points(pts, L, K, s),
                       oracle.sample_farthest_points_idx(pts, L, K, s))
    x = torch.randn(50, 4, generator=gen)
    f = torch.tensor([0, 10, 10, 37])
    assert torch.equal(R.packed_to_padded(x, f, 27), oracle.packed_to_padded_C(x, f, 27))
    pad = R.packed_to_padded(x, f, 27)
    assert torch.equal(R.padded_to_packed(pad, f, 50), oracle.padded_to_packed_C(pad, f, 50))


def test_oracle_threads_and_query_window(oracle):
    """The bounded-sample knobs used by bench.py's cpu_baseline leg do not change results."""
    gen = torch.Generator().manual_seed(6)
    p = torch.rand(2, 300, 3, generator=gen)
    full_i, full_d = oracle.knn_points_idx(p, p, None, None, 2, 16)
    win_i, win_d = oracle.knn_points_idx(p, p, None, None, 2, 16, q0=32, q1=96, threads=4)
    assert torch.equal(win_i[:, 32:96], full_i[:, 32:96]) and torch.equal(win_d[:, 32:96], full_d[:, 32:96])
    assert not win_i[:, :32].any() and not win_i[:, 96:].any()


def test_sample_pdf_oracle_vs_golden_and_ref(golden):
    """oracle_sample_pdf == the unmodified reference (golden vectors; oracle/_ref when present)."""
    from oracle import build_ref
    from oracle import oracle as O

    g = golden("sample_pdf_cases")
    for name, n_samples in (("a", 33), ("b", 128), ("c", 9), ("d", 17)):
        bins, w = g.t(f"{name}.bins"), g.t(f"{name}.weights")
        assert torch.equal(O.sample_pdf(bins, w, n_samples, det=True), g.t(f"{name}.det"))
        assert torch.equal(O.sample_pdf(bins, w, n_samples, u=g.t(f"{name}.u")), g.t(f"{name}.rand"))
    if build_ref.available():
        ref = build_ref.load()
        gen = torch.Generator().manual_seed(8)
        for B, n_bins, n_samples in ((8, 64, 50), (4, 3, 7), (1, 128, 300)):  # B: see make_golden.py
            bins = torch.sort(torch.randn(B, n_bins + 1, generator=gen), dim=-1).values
            w = torch.rand(B, n_bins, generator=gen) * (torch.rand(B, n_bins, generator=gen) > 0.2)
            u = torch.rand(B, n_samples, generator=gen)
            a, b = u.clone(), u.clone()
            ref.sample_pdf(bins, w, a, 1e-5)
            O.sample_pdf_(bins, w, b, 1e-5)
            assert torch.equal(a, b)


def test_oracle_point_covariances_vs_reference_golden(oracle, golden):
    """get_point_covariances (functions/utils.py:111-153) restated on the oracle's KNN + gather."""
    g = golden("utils_cases")
    for case, K in (("cov3", 8), ("cov3", 16), ("cov2", 6)):
        pts, L = g.t(f"{case}.points"), g.t(f"{case}.lengths")
        nn = oracle.knn_points(pts, pts, L, L, K=K, return_nn=True)[2]
        assert torch.equal(nn, g.t(f"{case}.K{K}.nn"))
        c = nn - nn.mean(2, keepdim=True)
        cov = (c.unsqueeze(4) * c.unsqueeze(3)).mean(2)
        assert torch.allclose(cov, g.t(f"{case}.K{K}.cov"), rtol=1e-6, atol=1e-8)


def test_oracle_non_finite_well_defined_cases(oracle, ref_module):
    """Where the reference's heap is well defined on non-finite data the oracle restates it: a NaN query
    keeps the first K points (push rule `size < K || dist < top`, knn_cpu.cpp:52), +inf distances are
    ordinary values."""
    gen = torch.Generator().manual_seed(5)
    p1, p2 = torch.rand(1, 6, 3, generator=gen), torch.rand(1, 40, 3, generator=gen)
    p1[0, 2, 1] = float("nan")
    p2[0, 3, 0] = float("inf")
    L1, L2 = torch.tensor([6]), torch.tensor([40])
    oi, od = oracle.knn_points_idx(p1, p2, L1, L2, 2, 5)
    assert oi[0, 2].tolist() == [0, 1, 2, 3, 4] and torch.isnan(od[0, 2]).all()
    oi40, od40 = oracle.knn_points_idx(p1, p2, L1, L2, 2, 40)
    assert oi40[0, 0, -1] == 3 and torch.isinf(od40[0, 0, -1])
    if ref_module is not None:
        ri, rd = ref_module.knn_points_idx(p1, p2, L1, L2, 2, 40, -1)
        assert torch.equal(ri, oi40)
        assert torch.equal(torch.isnan(rd), torch.isnan(od40)) and torch.equal(rd[~torch.isnan(rd)], od40[~torch.isnan(od40)])
