"""CPU: host-side logic that needs no GPU -- Pointclouds plumbing against the reference's
golden vectors, argument validation (same exceptions as the reference), alias package."""
import pytest
import torch


def _pc(golden):
    from pytorch3d_pointops_b200.structures import Pointclouds

    g = golden("pointclouds_cases")
    sizes = g.t("sizes").tolist()
    pts = [g.t(f"pts{i}") for i in range(len(sizes))]
    nrm = [g.t(f"nrm{i}") for i in range(len(sizes))]
    col = [g.t(f"col{i}") for i in range(len(sizes))]
    return g, Pointclouds(pts, features={"normals": nrm, "colors": col}), pts


def test_pointclouds_layouts_vs_reference(golden):
    g, pc, pts = _pc(golden)
    assert torch.equal(pc.points_padded(), g.t("points_padded"))
    assert torch.equal(pc.points_packed(), g.t("points_packed"))
    assert torch.equal(pc.features_padded()["normals"], g.t("normals_padded"))
    assert torch.equal(pc.features_packed()["colors"], g.t("colors_packed"))
    assert torch.equal(pc.num_points_per_cloud(), g.t("num_points_per_cloud"))
    assert torch.equal(pc.packed_to_cloud_idx(), g.t("packed_to_cloud_idx"))
    assert torch.equal(pc.cloud_to_packed_first_idx(), g.t("cloud_to_packed_first_idx"))
    assert torch.equal(pc.padded_to_packed_idx(), g.t("padded_to_packed_idx"))
    assert len(pc) == 4 and not pc.equisized and pc.valid.tolist() == [True, False, True, True]
    flat = pc.points_padded().reshape(-1, 3)
    assert torch.equal(flat[pc.padded_to_packed_idx()], pc.points_packed())


def test_pointclouds_batch_ops(golden):
    from pytorch3d_pointops_b200.structures import Pointclouds
    from pytorch3d_pointops_b200.structures.point_structure import (
        all_close, get_bounding_boxes, join_pointclouds_as_batch, join_pointclouds_as_scene,
        offset, scale, subsample)

    g, pc, pts = _pc(golden)
    sub = pc[[0, 2]]
    assert len(sub) == 2 and torch.equal(sub.points_list()[1], pts[2])
    assert torch.equal(pc[2:].points_list()[0], pts[2])
    assert torch.equal(pc[torch.tensor([True, False, False, True])].points_list()[1], pts[3])
    with pytest.raises(IndexError):
        pc[torch.tensor([0.5])]
    c = pc.clone()
    assert all_close(pc, c) and c.points_packed().data_ptr() != pc.points_packed().data_ptr()
    assert all_close(pc.detach(), pc)
    ext = pc.extend(2)
    assert len(ext) == 8 and torch.equal(ext.points_list()[5], pts[2])
    parts = pc.split([1, 3])
    assert [len(p) for p in parts] == [1, 3]
    pt, feats = pc.get_cloud(2)
    assert torch.equal(pt, pts[2]) and set(feats) == {"normals", "colors"}
    moved = offset(pc, torch.tensor([1.0, 2.0, 3.0]))
    assert torch.allclose(moved.points_packed(), pc.points_packed() + torch.tensor([1.0, 2.0, 3.0]))
    assert torch.allclose(moved.points_padded()[2, :12], pts[2] + torch.tensor([1.0, 2.0, 3.0]))
    sc = scale(pc, 2.0)
    assert torch.allclose(sc.points_packed(), pc.points_packed() * 2)
    joined = join_pointclouds_as_batch([pc, pc])
    assert len(joined) == 8 and set(joined.features_list()) == {"normals", "colors"}
    scene = join_pointclouds_as_scene(pc)
    assert len(scene) == 1 and scene.points_padded().shape == (1, 24, 3)
    nonempty = Pointclouds([pts[0], pts[2]])
    bb = get_bounding_boxes(nonempty)
    assert bb.shape == (2, 3, 2) and torch.equal(bb[0, :, 0], pts[0].min(0)[0])
    small = subsample(pc, 6)
    assert small.num_points_per_cloud().tolist() == [5, 0, 6, 6]
    assert small.features_list()["colors"][2].shape == (6, 4)
    assert subsample(pc, 100) is pc
    box = torch.tensor([[-10.0, -10, -10], [10, 10, 10]])
    assert pc.inside_box(box).all()
    per = box[None].expand(4, 2, 3).clone()
    per[0, 1] = -10.0
    inside = pc.inside_box(per)
    assert not inside[:5].any() and inside[5:].all()
    new = pc.update_padded(pc.points_padded() + 1.0)
    assert torch.equal(new.points_packed(), pc.points_packed() + 1.0)
    assert torch.equal(new.features_packed()["colors"], pc.features_packed()["colors"])
    newf = pc.update_padded(pc.points_padded(), {"normals": pc.features_padded()["normals"] * 0})
    assert set(newf.features_padded()) == {"normals"}


def test_pointclouds_constructor_errors():
    from pytorch3d_pointops_b200.structures import Pointclouds

    with pytest.raises(ValueError, match="incorrect dimensions"):
        Pointclouds(torch.zeros(2, 5, 4))
    with pytest.raises(ValueError, match="Px3 or empty"):
        Pointclouds([torch.zeros(5, 2)])
    with pytest.raises(ValueError, match="must be a dictionary"):
        Pointclouds([torch.zeros(5, 3)], features=[torch.zeros(5, 3)])
    with pytest.raises(ValueError, match="mismatched numbers"):
        Pointclouds([torch.zeros(5, 3)], features={"f": [torch.zeros(4, 3)]})
    with pytest.raises(ValueError, match="same number of channels"):
        Pointclouds([torch.zeros(5, 3), torch.zeros(2, 3)],
                    features={"f": [torch.zeros(5, 3), torch.zeros(2, 4)]})
    with pytest.raises(ValueError, match="Points must be either"):
        Pointclouds("nope")
    empty = Pointclouds([])
    assert len(empty) == 0 and empty.isempty() and empty.points_packed().shape == (0, 3)
    padded = Pointclouds(torch.rand(2, 4, 3), features={"c": torch.rand(2, 4, 2)})
    assert padded.equisized and padded.points_packed().shape == (8, 3)
    assert padded.features_list()["c"][1].shape == (4, 2)


def test_struct_utils_roundtrips():
    from pytorch3d_pointops_b200.structures import utils as U

    xs = [torch.randn(3, 2), torch.randn(0, 2), torch.randn(5, 2)]
    padded = U.list_to_padded(xs, (5, 2))
    assert padded.shape == (3, 5, 2) and torch.equal(padded[2], xs[2]) and not padded[1].any()
    assert torch.equal(U.list_to_padded([xs[0], xs[0]], equisized=True), torch.stack([xs[0]] * 2))
    with pytest.raises(ValueError, match="same number of dimensions"):
        U.list_to_padded([torch.zeros(2, 2), torch.zeros(2, 2, 2)])
    items = U.padded_to_list(padded, [3, 0, 5])
    assert [i.shape[0] for i in items] == [3, 0, 5]
    packed, counts, first, to_list = U.list_to_packed(xs)
    assert counts.tolist() == [3, 0, 5] and first.tolist() == [0, 3, 3]
    assert to_list.tolist() == [0, 0, 0, 2, 2, 2, 2, 2]
    assert [t.shape[0] for t in U.packed_to_list(packed, [3, 0, 5])] == [3, 0, 5]
    assert torch.equal(U.padded_to_packed(padded, split_size=[3, 0, 5]), packed)
    assert U.padded_to_packed(padded).shape == (15, 2)
    assert torch.equal(U.padded_to_packed(padded, pad_value=0.0), packed)
    with pytest.raises(ValueError, match="Only one of"):
        U.padded_to_packed(padded, split_size=[3, 0, 5], pad_value=0.0)
    with pytest.raises(ValueError, match="empty"):
        U.list_to_packed([])


def test_argument_validation_matches_reference():
    from pytorch3d_pointops_b200.functions import (ball_query, knn_gather, knn_points,
                                                   masked_gather, packed_to_padded,
                                                   padded_to_packed, sample_farthest_points, wmean)
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    a, b = torch.rand(2, 5, 3), torch.rand(3, 5, 3)
    with pytest.raises(ValueError, match="same batch dimension"):
        knn_points(a, b)
    with pytest.raises(ValueError, match="same point dimension"):
        knn_points(a, torch.rand(2, 5, 4))
    with pytest.raises(ValueError, match="same batch dimension"):
        ball_query(a, b)
    with pytest.raises(ValueError, match="same batch dimension"):
        knn_gather(a, torch.zeros(3, 5, 2, dtype=torch.int64))
    with pytest.raises(ValueError, match="same batch dimension"):
        masked_gather(a, torch.zeros(3, 2, dtype=torch.int64))
    with pytest.raises(ValueError, match="same batch dimension"):
        sample_farthest_points(a, lengths=torch.tensor([5]))
    with pytest.raises(ValueError, match="too large"):
        sample_farthest_points(a, lengths=torch.tensor([5, 6]))
    with pytest.raises(ValueError, match="same batch dimension"):
        sample_farthest_points(a, K=[1, 2, 3])
    with pytest.raises(ValueError, match="torch.float32"):
        packed_to_padded(torch.zeros(4, 2, dtype=torch.float64), torch.tensor([0]), 4)
    with pytest.raises(ValueError, match="torch.int64"):
        packed_to_padded(torch.zeros(4, 2), torch.tensor([0], dtype=torch.int32), 4)
    with pytest.raises(ValueError, match="has to be int"):
        packed_to_padded(torch.zeros(4, 2), torch.tensor([0]), 4.0)
    with pytest.raises(ValueError, match="first_idxs can only be 1-dimensional"):
        padded_to_packed(torch.zeros(1, 4), torch.tensor([[0]]), 4)
    with pytest.raises(ValueError, match="batch_reduction must be"):
        chamfer_distance(a, a, batch_reduction="max")
    with pytest.raises(ValueError, match="point_reduction must be"):
        chamfer_distance(a, a, point_reduction="min")
    with pytest.raises(ValueError, match="Batch reduction must be None"):
        chamfer_distance(a, a, point_reduction=None)
    with pytest.raises(ValueError, match="1 or 2 norm"):
        chamfer_distance(a, a, norm=3)
    with pytest.raises(ValueError, match='Features must be None if point_reduction is "max"'):
        chamfer_distance(a, a, point_reduction="max", feature_names=["n"])
    with pytest.raises(ValueError, match="Expected points to be of shape"):
        chamfer_distance(torch.rand(5, 3), a)
    with pytest.raises(ValueError, match="A length value was too long"):
        chamfer_distance(a, a, x_lengths=torch.tensor([5, 9]))
    with pytest.raises(ValueError, match="missing in x_features"):
        chamfer_distance(a, a, x_features={}, y_features={"n": a}, feature_names=["n"])
    with pytest.raises(ValueError, match="either"):
        chamfer_distance([a], a)
    w = wmean(torch.ones(2, 4, 3), torch.tensor([[1.0, 1, 0, 0], [1, 1, 1, 1]]))
    assert w.shape == (2, 1, 3) and torch.allclose(w, torch.ones(2, 1, 3))
    with pytest.raises(ValueError, match="not compatible"):
        wmean(torch.ones(2, 4, 3), torch.ones(2, 5))


def test_alias_package_is_a_drop_in():
    import pytorch3d_pointops
    import pytorch3d_pointops_b200
    from pytorch3d_pointops import _C  # noqa: F401
    from pytorch3d_pointops.functions import knn_points
    from pytorch3d_pointops.functions.chamfer import chamfer_distance  # noqa: F401
    from pytorch3d_pointops.structures import Pointclouds
    from pytorch3d_pointops.structures.point_structure import join_pointclouds_as_batch  # noqa: F401

    assert knn_points is pytorch3d_pointops_b200.functions.knn_points
    assert Pointclouds is pytorch3d_pointops_b200.structures.Pointclouds
    assert pytorch3d_pointops.__version__ == "0.7.8"
    for name in ("packed_to_padded", "padded_to_packed", "knn_check_version", "knn_points_idx",
                 "knn_points_backward", "ball_query", "sample_farthest_points"):
        assert callable(getattr(pytorch3d_pointops._C, name))  # ext.cpp:16-24


def test_wmean_vs_reference_golden(golden):
    """functions/utils.py:68-108 of the reference (pure torch on both sides)."""
    from pytorch3d_pointops_b200.functions.utils import wmean

    g = golden("utils_cases")
    x, w = g.t("wmean.x"), g.t("wmean.w")
    assert torch.equal(wmean(x), g.t("wmean.plain"))
    assert torch.equal(wmean(x, w), g.t("wmean.weighted"))
    assert torch.equal(wmean(x, w, keepdim=False), g.t("wmean.nokeep"))
    assert torch.equal(wmean(x, w, dim=(0, 1)), g.t("wmean.dim01"))
    assert torch.equal(wmean(x, w[:, :1]), g.t("wmean.bcast"))
    with pytest.raises(ValueError):
        wmean(x, torch.rand(4, 49))


def test_naive_fps_cpu_vs_reference_golden(golden):
    """The repo's sample_farthest_points_naive (reference functions/sample_farthest_points.py:99-197)
    is device-agnostic torch: on CPU tensors it must reproduce the reference's indices (big.idx was
    asserted equal to the reference's naive output at generation time)."""
    from pytorch3d_pointops_b200.functions.sample_farthest_points import sample_farthest_points_naive

    g = golden("fps_cases")
    sp, si = sample_farthest_points_naive(g.t("big.points"), K=200)
    assert torch.equal(si, g.t("big.idx")) and torch.equal(sp, g.t("big.sampled"))
    sp, si = sample_farthest_points_naive(g.t("ragged.points"), g.t("ragged.lengths"), g.t("ragged.K").tolist())
    assert torch.equal(si, g.t("ragged.idx")) and torch.equal(sp, g.t("ragged.sampled"))
    with pytest.raises(ValueError):
        sample_farthest_points_naive(g.t("ragged.points"), torch.tensor([1, 2]))


def test_chamfer_feature_shape_validation():
    """mis-shaped features must raise before any pointer reaches a kernel."""
    from pytorch3d_pointops_b200 import _C

    N, P1, P2 = 2, 5, 7
    ok_x, ok_y = torch.zeros(N, P1, 3), torch.zeros(N, P2, 3)
    _C._check_feature_shapes([ok_x], [ok_y], N, P1, P2)
    for bad_x, bad_y in [(torch.zeros(N, P1 - 1, 3), ok_y), (ok_x, torch.zeros(N, P2, 4)),
                         (torch.zeros(N + 1, P1, 3), ok_y), (ok_x, torch.zeros(N, P2 + 1, 3)),
                         (torch.zeros(N, P1), ok_y)]:
        with pytest.raises(ValueError):
            _C._check_feature_shapes([bad_x], [bad_y], N, P1, P2)
    with pytest.raises(ValueError):
        _C._check_feature_shapes([ok_x], [], N, P1, P2)
