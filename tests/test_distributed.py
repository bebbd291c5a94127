"""CPU, world_size 2, gloo: the host-side multi-GPU logic (sharding by cloud + the chamfer scalar
all-reduce).  The per-rank op is the CPU oracle here (the CUDA ops have no CPU path); on the GPU
box the same code runs the kernels -- see bench.py --gpus N."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def test_shard_bounds_balance():
    from pytorch3d_pointops_b200.distributed import shard_bounds

    assert shard_bounds([1] * 8, 2) == [(0, 4), (4, 8)]
    sizes = [hi - lo for lo, hi in shard_bounds([1] * 7, 4)]
    assert sum(sizes) == 7 and max(sizes) - min(sizes) <= 1
    sizes = [hi - lo for lo, hi in shard_bounds([1] * 2, 4)]
    assert sum(sizes) == 2 and max(sizes) <= 1
    assert shard_bounds([], 3) == [(0, 0)] * 3
    b = shard_bounds([10, 1, 1, 1, 1, 1, 1, 10], 2)
    assert b[0][0] == 0 and b[-1][1] == 8 and b[0][1] == b[1][0]
    costs = [10, 1, 1, 1, 1, 1, 1, 10]
    assert abs(sum(costs[b[0][0]:b[0][1]]) - sum(costs[b[1][0]:b[1][1]])) <= 10
    for world in (1, 2, 3, 5, 8):
        bb = shard_bounds([3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5], world)
        assert bb[0][0] == 0 and bb[-1][1] == 11
        assert all(bb[i][1] == bb[i + 1][0] for i in range(world - 1))


def _worker(rank, world, port, ret):
    import sys

    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from pytorch3d_pointops_b200.distributed import all_gather_clouds, chamfer_distance_sharded, my_slice

    g = torch.Generator().manual_seed(21)
    N, P = 5, 40
    x, y = torch.rand(N, P, 3, generator=g), torch.rand(N, P, 3, generator=g)
    xl, yl = torch.tensor([40, 13, 1, 30, 40]), torch.tensor([40, 40, 7, 3, 25])
    xn, yn = torch.randn(N, P, 3, generator=g), torch.randn(N, P, 3, generator=g)
    w = torch.tensor([1.0, 0.5, 0.0, 2.0, 1.5])
    ok = True
    for (wts, br, ncg) in ((None, "mean", N), (w, "mean", N), (None, "sum", N), (None, "mean", None)):
        xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        full, full_f = O.chamfer_distance(xr, yr, xl, yl, {"n": xn}, {"n": yn}, wts, br, "mean",
                                          feature_names=["n"])
        (full + full_f["n"]).backward()
        lo, hi = my_slice(N, costs=(xl * yl).tolist())
        xs, ys = x[lo:hi].clone().requires_grad_(True), y[lo:hi].clone().requires_grad_(True)
        loss, lf = chamfer_distance_sharded(
            xs, ys, xl[lo:hi], yl[lo:hi], {"n": xn[lo:hi]}, {"n": yn[lo:hi]},
            None if wts is None else wts[lo:hi], br, "mean", feature_names=["n"],
            n_clouds_global=ncg, _local_fn=O.chamfer_distance)
        (loss + lf["n"]).backward()
        ok &= torch.allclose(loss.detach(), full.detach(), rtol=1e-5, atol=1e-7)
        ok &= torch.allclose(lf["n"].detach(), full_f["n"].detach(), rtol=1e-5, atol=1e-7)
        ok &= torch.allclose(xs.grad, xr.grad[lo:hi], rtol=1e-5, atol=1e-8)
        ok &= torch.allclose(ys.grad, yr.grad[lo:hi], rtol=1e-5, atol=1e-8)
    # gather of per-cloud outputs with unequal shards
    lo, hi = my_slice(N)
    idx, _ = O.knn_points_idx(x[lo:hi], y[lo:hi], xl[lo:hi], yl[lo:hi], 2, 3)
    gathered = all_gather_clouds(idx)
    want, _ = O.knn_points_idx(x, y, xl, yl, 2, 3)
    ok &= torch.equal(gathered, want)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_chamfer_and_gather_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        port = 29500 + (os.getpid() % 500)
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}


def _nccl_worker(rank, world, port, ret):
    """One rank = one GPU: the CUDA chamfer on this rank's clouds + the NCCL all-reduce; the global
    loss must equal the single-GPU loss on the whole batch, the gradients the matching slices."""
    import sys

    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pytorch3d_pointops_b200.distributed import all_gather_clouds, chamfer_distance_sharded, my_slice
    from pytorch3d_pointops_b200.functions import knn_points
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    g = torch.Generator().manual_seed(33)
    N, P = 6, 3000
    x, y = torch.rand(N, P, 3, generator=g).to(dev), torch.rand(N, P, 3, generator=g).to(dev)
    xl = torch.tensor([3000, 1300, 1, 2999, 2048, 77], device=dev)
    yl = torch.tensor([3000, 3000, 700, 3, 2500, 1024], device=dev)
    xn, yn = torch.randn(N, P, 3, generator=g).to(dev), torch.randn(N, P, 3, generator=g).to(dev)
    ok = True
    for br in ("mean", "sum"):
        xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        full, full_f = chamfer_distance(xr, yr, xl, yl, {"n": xn}, {"n": yn}, batch_reduction=br, feature_names=["n"])
        (full + full_f["n"]).backward()
        lo, hi = my_slice(N, costs=(xl * yl).tolist())
        xs, ys = x[lo:hi].clone().requires_grad_(True), y[lo:hi].clone().requires_grad_(True)
        loss, lf = chamfer_distance_sharded(xs, ys, xl[lo:hi], yl[lo:hi], {"n": xn[lo:hi]}, {"n": yn[lo:hi]},
                                            batch_reduction=br, feature_names=["n"], n_clouds_global=N)
        (loss + lf["n"]).backward()
        ok &= bool(torch.allclose(loss.detach(), full.detach(), rtol=1e-5))
        ok &= bool(torch.allclose(lf["n"].detach(), full_f["n"].detach(), rtol=1e-5))
        scale = float(xr.grad.abs().max())
        ok &= bool(torch.allclose(xs.grad, xr.grad[lo:hi], rtol=1e-5, atol=1e-5 * scale))
        ok &= bool(torch.allclose(ys.grad, yr.grad[lo:hi], rtol=1e-5, atol=1e-5 * scale))
    lo, hi = my_slice(N)
    idx = knn_points(x[lo:hi], y[lo:hi], xl[lo:hi], yl[lo:hi], K=3).idx
    ok &= bool(torch.equal(all_gather_clouds(idx), knn_points(x, y, xl, yl, K=3).idx))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_chamfer_nccl_two_gpus():
    """The path's one collective on real NCCL hardware (needs >= 2 visible GPUs: `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        port = 29500 + (os.getpid() % 500)
        procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}
