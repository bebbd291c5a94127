"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py

The reference's python package is imported from /root/reference (read-only) and its
`pytorch3d_pointops._C` is satisfied by oracle/_ref/_C_ref*.so -- the reference's own CPU
sources compiled by oracle/build_ref.py.  Every fixture stores the inputs and the outputs of
the reference's public API (functions/*.py, structures/), so the tests need neither
/root/reference nor torch-RNG reproducibility.  /root/reference does not exist on the GPU
box; only the committed .npz files travel.
"""
import importlib.util
import os
import sys
import sysconfig

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def load_reference():
    so = os.path.join(REPO, "oracle", "_ref", "_C_ref" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.isfile(so):
        sys.path.insert(0, REPO)
        from oracle import build_ref

        build_ref.build()
        sys.path.pop(0)
    spec = importlib.util.spec_from_file_location("_C_ref", so)
    cref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cref)
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    sys.path.insert(0, REF)
    sys.modules["pytorch3d_pointops._C"] = cref
    import pytorch3d_pointops  # noqa: F401  (the reference package)

    assert pytorch3d_pointops.__file__.startswith(REF)
    pytorch3d_pointops._C = cref
    return cref


def npy(t):
    if t is None:
        return np.zeros((0,), np.float32)
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (npy(v) if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def utils_cases():
    """get_point_covariances / wmean (functions/utils.py:68-153 of the reference) on the reference's
    own CPU KNN: ragged clouds, K above the shortest cloud, D = 2 and 3, weighted means."""
    from pytorch3d_pointops.functions.utils import get_point_covariances, wmean

    g = torch.Generator().manual_seed(31)
    cases = {}
    pts = torch.rand(3, 120, 3, generator=g)
    L = torch.tensor([120, 5, 77])
    for K in (8, 16):  # K = 8 exceeds the 5-point cloud: its slots k >= 5 gather zeros (knn_gather)
        cov, nn = get_point_covariances(pts, L, K)
        cases.update({f"cov3.K{K}.cov": cov, f"cov3.K{K}.nn": nn})
    cases.update({"cov3.points": pts, "cov3.lengths": L})
    pts2 = torch.randn(2, 64, 2, generator=g)
    L2 = torch.tensor([64, 33])
    cov, nn = get_point_covariances(pts2, L2, 6)
    cases.update({"cov2.points": pts2, "cov2.lengths": L2, "cov2.K6.cov": cov, "cov2.K6.nn": nn})
    x = torch.randn(4, 50, 3, generator=g)
    w = torch.rand(4, 50, generator=g)
    w[1] = 0.0  # an all-zero weight row: the eps clamp decides
    cases.update({"wmean.x": x, "wmean.w": w, "wmean.plain": wmean(x), "wmean.weighted": wmean(x, w),
                  "wmean.nokeep": wmean(x, w, keepdim=False), "wmean.dim01": wmean(x, w, dim=(0, 1)),
                  "wmean.bcast": wmean(x, w[:, :1])})
    save("utils_cases", **cases)


def main():
    cref = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "utils":  # only the fixtures added in round 2
        utils_cases()
        return
    from pytorch3d_pointops.functions import (ball_query, knn_gather, knn_points,
                                              masked_gather, packed_to_padded,
                                              padded_to_packed, sample_farthest_points)
    from pytorch3d_pointops.functions.chamfer import chamfer_distance
    from pytorch3d_pointops.functions.sample_farthest_points import sample_farthest_points_naive
    from pytorch3d_pointops.structures import Pointclouds

    # ---------------------------------------------------------------- README config (C1)
    torch.manual_seed(0)
    pts = [torch.randn(1000, 3), torch.randn(800, 3)]
    pc = Pointclouds(points=pts)
    X, L = pc.points_padded(), pc.num_points_per_cloud()
    out = knn_points(X, X, lengths1=L, lengths2=L, K=8)
    assert abs(float(out.dists.sum()) - 2895.012695) < 1e-2, float(out.dists.sum())
    save("knn_readme", p0=pts[0], p1=pts[1], padded=X, lengths=L, dists=out.dists,
         idx=out.idx.to(torch.int32))

    # ---------------------------------------------------------------- KNN cases
    cases = {}
    g = torch.Generator().manual_seed(11)
    specs = [  # name, N, P1, P2, D, K, norm, ragged
        ("d3k1", 3, 70, 90, 3, 1, 2, True),
        ("d3k16", 2, 300, 257, 3, 16, 2, True),
        ("d3k32_l1", 2, 120, 150, 3, 32, 1, True),
        ("d3k40", 2, 64, 100, 3, 40, 2, False),
        ("d2k5", 2, 50, 60, 2, 5, 2, True),
        ("d5k7", 2, 50, 60, 5, 7, 2, True),
        ("d33k4", 2, 40, 45, 33, 4, 2, True),
        ("d128k16", 1, 48, 200, 128, 16, 2, False),
        ("klen", 3, 20, 6, 3, 9, 2, True),  # K > lengths2
    ]
    for name, N, P1, P2, D, K, norm, ragged in specs:
        p1 = torch.randn(N, P1, D, generator=g)
        p2 = torch.randn(N, P2, D, generator=g)
        if ragged:
            l1 = torch.randint(0, P1 + 1, (N,), generator=g)
            l2 = torch.randint(0, P2 + 1, (N,), generator=g)
            l1[0], l2[0] = P1, P2
            if name == "klen":
                l2 = torch.tensor([6, 3, 0])
        else:
            l1 = torch.full((N,), P1, dtype=torch.int64)
            l2 = torch.full((N,), P2, dtype=torch.int64)
        p1r = p1.clone().requires_grad_(True)
        p2r = p2.clone().requires_grad_(True)
        # canonical contract = raw CPU kernel order (return_sorted=False), SURVEY 2.2
        res = knn_points(p1r, p2r, l1, l2, norm=norm, K=K, return_nn=True, return_sorted=False)
        gd = torch.randn(res.dists.shape, generator=g)
        gn = torch.randn(res.knn.shape, generator=g)
        ((res.dists * gd).sum() + (res.knn * gn).sum()).backward()
        cases.update({f"{name}.p1": p1, f"{name}.p2": p2, f"{name}.l1": l1, f"{name}.l2": l2,
                      f"{name}.K": K, f"{name}.norm": norm, f"{name}.dists": res.dists,
                      f"{name}.idx": res.idx.to(torch.int32), f"{name}.knn": res.knn,
                      f"{name}.gd": gd, f"{name}.gn": gn, f"{name}.grad_p1": p1r.grad,
                      f"{name}.grad_p2": p2r.grad})
    # exact ties: 5^3 integer grid, every K (canonical order), and the SURVEY tie vector
    grid = torch.stack(torch.meshgrid(*[torch.arange(5.0)] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    for K in (1, 3, 7, 16, 17, 32, 33):
        i, d = cref.knn_points_idx(grid, grid, torch.tensor([125]), torch.tensor([125]), 2, K, -1)
        cases[f"grid.K{K}.idx"] = i.to(torch.int32)
        cases[f"grid.K{K}.dists"] = d
    cases["grid.p"] = grid
    tie_p1 = torch.tensor([[[0.0, 0, 0]]])
    tie_p2 = torch.tensor([[[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0]]])
    for K in (3, 4, 5):
        r = knn_points(tie_p1, tie_p2, K=K)
        cases[f"tie.K{K}.idx"] = r.idx.to(torch.int32)
    cases["tie.p1"], cases["tie.p2"] = tie_p1, tie_p2
    save("knn_cases", **cases)

    # ---------------------------------------------------------------- ball query
    cases = {}
    g = torch.Generator().manual_seed(12)
    for name, N, P1, P2, D, K, radius in [("a", 2, 200, 300, 3, 8, 0.6),
                                          ("b", 2, 100, 150, 3, 32, 1.5),
                                          ("c", 2, 64, 64, 4, 5, 0.9),
                                          ("none", 1, 16, 16, 3, 4, 1e-3)]:
        p1 = torch.randn(N, P1, D, generator=g)
        p2 = torch.randn(N, P2, D, generator=g)
        l1 = torch.randint(1, P1 + 1, (N,), generator=g)
        l2 = torch.randint(1, P2 + 1, (N,), generator=g)
        p1r = p1.clone().requires_grad_(True)
        p2r = p2.clone().requires_grad_(True)
        res = ball_query(p1r, p2r, l1, l2, K=K, radius=radius, return_nn=True)
        gd = torch.randn(res.dists.shape, generator=g)
        gn = torch.randn(res.knn.shape, generator=g)
        ((res.dists * gd).sum() + (res.knn * gn).sum()).backward()
        cases.update({f"{name}.p1": p1, f"{name}.p2": p2, f"{name}.l1": l1, f"{name}.l2": l2,
                      f"{name}.K": K, f"{name}.radius": radius, f"{name}.dists": res.dists,
                      f"{name}.idx": res.idx.to(torch.int32), f"{name}.knn": res.knn,
                      f"{name}.gd": gd, f"{name}.gn": gn, f"{name}.grad_p1": p1r.grad,
                      f"{name}.grad_p2": p2r.grad})
    lin = torch.linspace(0, 1, 6)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3)
    res = ball_query(grid, grid, K=30, radius=0.25)
    cases.update({"grid.p": grid, "grid.idx": res.idx.to(torch.int32), "grid.dists": res.dists})
    save("ball_query_cases", **cases)

    # ---------------------------------------------------------------- FPS
    cases = {}
    torch.manual_seed(1)
    pts = torch.rand(3, 50, 3)
    L = torch.tensor([50, 7, 20])
    K = [10, 10, 30]
    sp, si = sample_farthest_points(pts, L, K)
    assert si[1].tolist()[:8] == [0, 1, 2, 6, 4, 5, 3, -1], si[1].tolist()
    cases.update({"ragged.points": pts, "ragged.lengths": L, "ragged.K": torch.tensor(K),
                  "ragged.idx": si.to(torch.int32), "ragged.sampled": sp})
    torch.manual_seed(456)
    pts = torch.randn(2, 2000, 3)
    sp, si = sample_farthest_points(pts, K=200)
    sp2, si2 = sample_farthest_points_naive(pts, K=200)
    assert torch.equal(si, si2)
    cases.update({"big.points": pts, "big.idx": si.to(torch.int32), "big.sampled": sp})
    pts = torch.zeros(1, 5, 3)
    _, si = sample_farthest_points(pts, K=4)
    cases.update({"dup.points": pts, "dup.idx": si.to(torch.int32)})
    g = torch.Generator().manual_seed(13)
    pts = torch.rand(4, 300, 3, generator=g)
    L = torch.tensor([300, 1, 150, 299])
    Kt = torch.tensor([64, 5, 150, 2])
    start = torch.tensor([17, 0, 149, 5])
    si = cref.sample_farthest_points(pts, L, Kt, start)
    cases.update({"start.points": pts, "start.lengths": L, "start.K": Kt, "start.start": start,
                  "start.idx": si.to(torch.int32)})
    pts = torch.rand(2, 90, 6, generator=g)
    sp, si = sample_farthest_points(pts, K=33)
    cases.update({"d6.points": pts, "d6.idx": si.to(torch.int32)})
    save("fps_cases", **cases)

    # ---------------------------------------------------------------- packed <-> padded
    cases = {}
    torch.manual_seed(42)
    sizes = [100, 500, 2000, 25]
    first = torch.tensor([0] + list(np.cumsum(sizes)[:-1]), dtype=torch.int64)
    for D, shape in [(3, (3,)), (1, ()), (6, (2, 3))]:
        packed = torch.randn(sum(sizes), *shape, requires_grad=True)
        padded = packed_to_padded(packed, first, max(sizes))
        gpad = torch.randn(padded.shape)
        (padded * gpad).sum().backward()
        back = padded_to_packed(padded.detach(), first, sum(sizes))
        assert torch.equal(back, packed.detach())
        cases.update({f"D{D}.packed": packed, f"D{D}.padded": padded, f"D{D}.gpad": gpad,
                      f"D{D}.grad_packed": packed.grad})
    cases["first"] = first
    cases["sizes"] = torch.tensor(sizes)
    save("packed_padded_cases", **cases)

    # ---------------------------------------------------------------- gathers
    cases = {}
    g = torch.Generator().manual_seed(14)
    x = torch.randn(3, 40, 7, generator=g)
    idx = torch.randint(0, 40, (3, 25, 6), generator=g)
    lens = torch.tensor([40, 4, 0])
    cases.update({"kg.x": x, "kg.idx": idx.to(torch.int32), "kg.lengths": lens,
                  "kg.out": knn_gather(x, idx, lens), "kg.out_full": knn_gather(x, idx)})
    idx3 = idx.clone()
    idx3[torch.rand(idx3.shape, generator=g) < 0.3] = -1
    idx2 = torch.randint(-1, 40, (3, 9), generator=g)
    cases.update({"mg.idx3": idx3.to(torch.int32), "mg.out3": masked_gather(x, idx3),
                  "mg.idx2": idx2.to(torch.int32), "mg.out2": masked_gather(x, idx2)})
    save("gather_cases", **cases)

    # ---------------------------------------------------------------- chamfer
    cases = {}
    g = torch.Generator().manual_seed(15)
    N, P1, P2 = 4, 150, 130
    x = torch.rand(N, P1, 3, generator=g)
    y = torch.rand(N, P2, 3, generator=g)
    xl = torch.tensor([150, 90, 1, 120])
    yl = torch.tensor([130, 130, 7, 64])
    xn = torch.nn.functional.normalize(torch.randn(N, P1, 3, generator=g), dim=-1)
    yn = torch.nn.functional.normalize(torch.randn(N, P2, 3, generator=g), dim=-1)
    xc = torch.rand(N, P1, 4, generator=g)
    yc = torch.rand(N, P2, 4, generator=g)
    w = torch.tensor([1.0, 0.5, 0.0, 2.0])
    cases.update(dict(x=x, y=y, xl=xl, yl=yl, xn=xn, yn=yn, xc=xc, yc=yc, w=w))
    variants = []
    for br in ("mean", "sum", None):
        for pr in ("mean", "sum", "max", None):
            if pr is None and br is not None:
                continue
            for feats in (True, False):
                if pr == "max" and feats:
                    continue
                variants.append(dict(br=br, pr=pr, feats=feats, norm=2, sd=False, w=False,
                                     abs=True, ragged=True))
    variants += [
        dict(br="mean", pr="mean", feats=True, norm=1, sd=False, w=True, abs=False, ragged=True),
        dict(br="mean", pr="mean", feats=True, norm=2, sd=True, w=True, abs=True, ragged=True),
        dict(br="sum", pr="mean", feats=False, norm=2, sd=False, w=False, abs=True, ragged=False),
        dict(br=None, pr=None, feats=True, norm=2, sd=True, w=False, abs=True, ragged=True),
        dict(br=None, pr="max", feats=False, norm=2, sd=True, w=True, abs=True, ragged=True),
    ]
    meta = []
    for vi, v in enumerate(variants):
        ts = [t.clone().requires_grad_(True) for t in (x, y, xn, yn, xc, yc)]
        xr, yr, xnr, ynr, xcr, ycr = ts
        kw = dict(batch_reduction=v["br"], point_reduction=v["pr"], norm=v["norm"],
                  single_directional=v["sd"], abs_cosine=v["abs"])
        if v["ragged"]:
            kw.update(x_lengths=xl, y_lengths=yl)
        if v["w"]:
            kw["weights"] = w
        if v["feats"]:
            kw.update(x_features={"normals": xnr, "colors": xcr},
                      y_features={"normals": ynr, "colors": ycr},
                      feature_names=["normals", "colors"])
        loss, lf = chamfer_distance(xr, yr, **kw)
        flat = []

        def walk(o):
            if o is None:
                return
            if torch.is_tensor(o):
                flat.append(o)
            elif isinstance(o, dict):
                for k in sorted(o):
                    walk(o[k])
            else:
                for e in o:
                    walk(e)

        walk(loss)
        walk(lf)
        total = sum((t * (i + 1)).sum() for i, t in enumerate(flat))
        total.backward()
        for i, t in enumerate(flat):
            cases[f"v{vi}.out{i}"] = t
        cases[f"v{vi}.nout"] = len(flat)
        for nm, t in zip(("x", "y", "xn", "yn", "xc", "yc"), ts):
            cases[f"v{vi}.g_{nm}"] = t.grad if t.grad is not None else torch.zeros(0)
        meta.append(repr(v))
    cases["variants"] = np.array(meta)
    # Pointclouds input with feature dicts == tensor input (examples/chamfer_loss.py:13-89)
    pcx = Pointclouds([x[i, : xl[i]] for i in range(N)],
                      features={"normals": [xn[i, : xl[i]] for i in range(N)]})
    pcy = Pointclouds([y[i, : yl[i]] for i in range(N)],
                      features={"normals": [yn[i, : yl[i]] for i in range(N)]})
    loss, lf = chamfer_distance(pcx, pcy, feature_names=["normals"])
    cases["pc.loss"], cases["pc.normals"] = loss, lf["normals"]
    save("chamfer_cases", **cases)

    # ---------------------------------------------------------------- Pointclouds plumbing
    cases = {}
    g = torch.Generator().manual_seed(16)
    sizes = [5, 0, 12, 7]
    pts = [torch.randn(s, 3, generator=g) for s in sizes]
    nrm = [torch.randn(s, 3, generator=g) for s in sizes]
    col = [torch.rand(s, 4, generator=g) for s in sizes]
    pc = Pointclouds(pts, features={"normals": nrm, "colors": col})
    for i, s in enumerate(sizes):
        cases[f"pts{i}"], cases[f"nrm{i}"], cases[f"col{i}"] = pts[i], nrm[i], col[i]
    cases.update({
        "sizes": torch.tensor(sizes),
        "points_padded": pc.points_padded(), "points_packed": pc.points_packed(),
        "normals_padded": pc.features_padded()["normals"],
        "colors_packed": pc.features_packed()["colors"],
        "num_points_per_cloud": pc.num_points_per_cloud(),
        "packed_to_cloud_idx": pc.packed_to_cloud_idx(),
        "cloud_to_packed_first_idx": pc.cloud_to_packed_first_idx(),
        "padded_to_packed_idx": pc.padded_to_packed_idx(),
    })
    save("pointclouds_cases", **cases)

    # ---- sample_pdf (functions/sample_pdf.py:14-66 on the reference's CPU kernel) -----------------
    from pytorch3d_pointops.functions.sample_pdf import sample_pdf

    g = torch.Generator().manual_seed(21)
    cases = {}
    # (batch sizes chosen so that the reference's 4-thread split stays in bounds: with B = 5 its
    #  third worker is handed rows [4, 6) -- sample_pdf_cpu.cpp:121-135 -- and corrupts the heap)
    for name, (shape, n_bins, n_samples) in {"a": ((4,), 64, 33), "b": ((3, 4), 7, 128), "c": ((1,), 1, 9),
                                             "d": ((2,), 200, 17)}.items():
        edges = torch.sort(torch.rand(shape + (n_bins + 1,), generator=g) * 4 - 1, dim=-1).values
        w = torch.rand(shape + (n_bins,), generator=g)
        if name == "b":
            w[..., 2] = 0.0          # empty bins
            w[0, 0] = 0.0            # an all-empty row
        cases[f"{name}.bins"], cases[f"{name}.weights"] = edges, w
        cases[f"{name}.det"] = sample_pdf(edges, w, n_samples, det=True)
        u = torch.rand(shape + (n_samples,), generator=g)
        out = u.clone()
        cref.sample_pdf(edges.reshape(-1, n_bins + 1), w.reshape(-1, n_bins), out.view(-1, n_samples), 1e-5)
        cases[f"{name}.u"], cases[f"{name}.rand"] = u, out
    save("sample_pdf_cases", **cases)
    utils_cases()


if __name__ == "__main__":
    main()
