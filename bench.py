#!/usr/bin/env python
"""bench.py -- headline benchmark of the nearest-neighbour hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "KNN queries/sec (BxP1, K=16, D=3) and chamfer pairs/sec"):
  knn_T    : self-KNN, B=32 clouds x P=16384 points, D=3, K=16, fp32, uniform rand (seed 0+rank)
  chamfer  : chamfer_distance fwd+bwd, B=32, P<=8192 ragged, normals+colors (reported under
             "secondary", same protocol)
A step = one pass of the hot path over one batch.  Weak scaling: every rank owns its own batch
(clouds shard by batch index; no data-path collective), value = queries of all ranks / max time.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

B, P, K_NN, D = 32, 16384, 16, 3
METRIC = "knn_queries_per_sec"
UNIT = "queries/s"
WORKLOAD = f"knn_points self-KNN B={B} P={P} D={D} K={K_NN} fp32 uniform (north_star target shape)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------
# clocks sampler (NVML in-process; nvidia-smi fallback)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period: float = 0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._nvml is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return self._smi_fallback()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": "nvml"}

    def _smi_fallback(self):
        import subprocess

        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                 "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b, c = [x.strip() for x in out.strip().split(",")[:3]]
            return {"sm_mhz": int(a), "sm_max_mhz": int(b), "reasons": [c], "samples": 1,
                    "source": "nvidia-smi (after the timed region)"}
        except Exception as e:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": f"unavailable: {e}"}


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
def make_knn_inputs(rank: int):
    g = torch.Generator().manual_seed(0 + rank)
    p = torch.rand(B, P, D, generator=g)
    lengths = torch.full((B,), P, dtype=torch.int64)
    return p, lengths


def make_chamfer_inputs(rank: int):
    g = torch.Generator().manual_seed(1 + 1000 * rank)
    N, Pc = 32, 8192
    x, y = torch.rand(N, Pc, 3, generator=g), torch.rand(N, Pc, 3, generator=g)
    xl = torch.randint(4096, Pc + 1, (N,), generator=g)
    yl = torch.randint(4096, Pc + 1, (N,), generator=g)
    xn = torch.nn.functional.normalize(torch.randn(N, Pc, 3, generator=g), dim=-1)
    yn = torch.nn.functional.normalize(torch.randn(N, Pc, 3, generator=g), dim=-1)
    xc, yc = torch.rand(N, Pc, 3, generator=g), torch.rand(N, Pc, 3, generator=g)
    return dict(x=x, y=y, xl=xl, yl=yl, xn=xn, yn=yn, xc=xc, yc=yc)


def ncu_traffic_bytes(summary_name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of the
    kernel (profiles/<summary_name>), per launch, or None."""
    path = os.path.join(REPO, "profiles", summary_name)
    if not os.path.isfile(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    with open(path) as fh:
        for ln in fh:
            ln = ln.strip()
            if ln.startswith(("dram__bytes_read.sum =", "dram__bytes_write.sum =")):
                val, u = ln.split("=")[1].split()[:2]
                total += float(val) * unit.get(u, 1.0)
                seen += 1
    return total if seen == 2 else None


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU reference / baseline
# --------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One worker = one host core running the reference's own CPU loop on a query slice."""
    (cloud, q0, q1, use_ref) = args
    import torch as _t

    _t.set_num_threads(1)
    p, lengths = _WORKER_STATE["p"], _WORKER_STATE["lengths"]
    p1 = p[cloud:cloud + 1, q0:q1].contiguous()
    p2 = p[cloud:cloud + 1]
    l1 = _t.tensor([q1 - q0])
    l2 = lengths[cloud:cloud + 1]
    t0 = time.perf_counter()
    if use_ref:
        _WORKER_STATE["ref"].knn_points_idx(p1, p2, l1, l2, 2, K_NN, -1)
    else:
        _WORKER_STATE["oracle"].knn_points_idx(p1, p2, l1, l2, 2, K_NN)
    return q1 - q0, time.perf_counter() - t0


_WORKER_STATE = {}


def _init_worker_state(p, lengths):
    from oracle import build_ref

    _WORKER_STATE["p"], _WORKER_STATE["lengths"] = p, lengths
    use_ref = build_ref.available()
    if use_ref:
        _WORKER_STATE["ref"] = build_ref.load()
    else:
        from oracle import oracle as O

        O.build()
        _WORKER_STATE["oracle"] = O
    return use_ref


def cpu_reference_rate(p, lengths, budget_s: float, workers: int):
    """Time the reference's CPU KNN (oracle/_ref when present, else the oracle port) on a bounded
    sample: each worker gets a slice of queries of its own cloud against the full P2.
    Returns (queries/s aggregate, kind, sample description, seconds)."""
    import multiprocessing as mp

    use_ref = _init_worker_state(p, lengths)
    # calibrate on one core: 128 queries of cloud 0
    n, dt = _ref_worker((0, 0, 128, use_ref))
    rate1 = n / max(dt, 1e-6)
    per_worker = int(max(64, rate1 * budget_s))          # queries one core finishes in the budget
    # split into whole-cloud-sized jobs: worker w walks clouds w, w+workers, ...
    jobs = []
    for w in range(workers):
        left, c = per_worker, w
        while left > 0:
            q = min(left, P)
            jobs.append((c % B, 0, q, use_ref))
            left -= q
            c += workers
    t0 = time.perf_counter()
    if workers == 1:
        res = [_ref_worker(j) for j in jobs]
    else:
        ctx = mp.get_context("fork")  # workers inherit the tensors and the loaded module
        with ctx.Pool(workers) as pool:
            res = pool.map(_ref_worker, jobs, chunksize=max(1, len(jobs) // workers))
    wall = time.perf_counter() - t0
    total_q = sum(r[0] for r in res)
    kind = "reference" if use_ref else "port"
    sample = (f"{workers} worker process(es) x {per_worker} queries (whole clouds of the batch, each query "
              f"against the full P2={P}; K={K_NN}, D={D}); the reference's native loop is single-threaded")
    return total_q / wall, kind, sample, wall


# --------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    p, lengths = make_knn_inputs(0)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_step_budget = max(0.25, min(4.0, 150.0 / (steps + warm)))
    vals, secs = [], []
    kind = sample = None
    for i in range(warm + steps):
        rate, kind, sample, wall = cpu_reference_rate(p, lengths, per_step_budget, workers)
        if i >= warm:
            vals.append(rate)
            secs.append(wall)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from pytorch3d_pointops_b200 import _C, _lib
    from pytorch3d_pointops_b200.functions import knn_points
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        # NCCL prints its version banner on STDOUT when the communicator is created (NCCL_DEBUG=VERSION
        # on the GPU boxes); stdout must carry the one JSON line only, so the banner goes to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    steps, warm = max(1, args.steps), max(3, args.warmup)
    p_host, len_host = make_knn_inputs(rank)
    p_pin = p_host.pin_memory()
    p_dev = p_host.to(dev)
    len_dev = len_host.to(dev)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    queries_per_step = int(len_host.sum())
    pairs_per_step = int((len_host * len_host).sum())

    def step_resident():
        return _C.knn_points_idx(p_dev, p_dev, len_dev, len_dev, 2, K_NN, -1)

    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize(dev)

    # ---- timed region: device-resident inputs, L2 flushed between steps -------------------------
    sampler = ClockSampler(local_rank)
    lib.pops_profile_reset()
    lib.pops_profile_enable(1)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    launches0 = _lib.launch_count()
    sampler.start()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        step_resident()
        b.record()
    torch.cuda.synchronize(dev)
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    lib.pops_profile_enable(0)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = queries_per_step * world * steps / (total_ms * 1e-3)

    import ctypes

    def kernel_ms(name):
        n_, ms_ = ctypes.c_int64(0), ctypes.c_double(0.0)
        lib.pops_profile_read(name, ctypes.byref(n_), ctypes.byref(ms_))
        return ms_.value / max(1, n_.value), int(n_.value)

    scan_ms, scan_launches = kernel_ms(b"knn_scan")
    n_l = ctypes.c_int64(scan_launches)
    lib.pops_profile_reset()

    # ---- what the pruned search executes: block counters of one untimed step, and the same kernel
    #      with pruning switched off (every block visited, same order) as the brute-force reference
    lib.pops_set_option(b"knn_stats", 1)
    stats = (ctypes.c_ulonglong * 8)()
    lib.pops_knn_debug_stats(stats)
    step_resident()
    lib.pops_knn_debug_stats(stats)
    lib.pops_set_option(b"knn_stats", 0)
    blocks_scanned, warps = int(stats[1]), max(1, int(stats[5]))
    executed_fraction = (blocks_scanned / warps) * 64.0 / P
    lib.pops_set_option(b"knn_prune", 0)
    lib.pops_set_option(b"knn_q", 4)  # the best brute-force shape of this kernel: a point read serves 4 queries
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize(dev)
    lib.pops_profile_enable(1)
    for _ in range(5):
        flush.zero_()
        step_resident()
    torch.cuda.synchronize(dev)
    lib.pops_profile_enable(0)
    brute_ms, _ = kernel_ms(b"knn_scan")
    lib.pops_profile_reset()
    lib.pops_set_option(b"knn_prune", 1)
    lib.pops_set_option(b"knn_q", 0)

    # ---- end to end: pinned host inputs -> H2D -> knn_points -> D2H of (dists, idx) ---------------
    out_d_pin = torch.empty((B, P, K_NN), dtype=torch.float32).pin_memory()
    out_i_pin = torch.empty((B, P, K_NN), dtype=torch.int64).pin_memory()
    len_pin = len_host.pin_memory()

    def step_e2e_serial():
        pd = p_pin.to(dev, non_blocking=True)
        ld = len_pin.to(dev, non_blocking=True)
        r = knn_points(pd, pd, ld, ld, K=K_NN)
        out_d_pin.copy_(r.dists, non_blocking=True)
        out_i_pin.copy_(r.idx, non_blocking=True)

    from pytorch3d_pointops_b200.host import HostKnn

    host_knn = HostKnn(B, P, P, D, K_NN, dev, slices=8)

    def step_e2e():
        # host-in / host-out API: the first of 8 slices on its own, then one H2D + one pre-pass for the rest, searched one slice
        # after the other while the D2H of the previous slice's results runs (one CUDA graph)
        host_knn(p_pin, None, len_pin)

    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_steps = max(3, min(steps, 20))
    e_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    barrier()
    for a, b in e_evs:
        flush.zero_()
        a.record()
        step_e2e()
        b.record()
    torch.cuda.synchronize(dev)
    barrier()
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in e_evs))
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = queries_per_step * world * e2e_steps / (e2e_ms * 1e-3)
    h2d = p_pin.numel() * 4 + len_pin.numel() * 8
    d2h = out_d_pin.numel() * 4 + out_i_pin.numel() * 8
    # the plain call sequence a user of the reference API writes (no overlap), for comparison
    for _ in range(2):
        step_e2e_serial()
    torch.cuda.synchronize(dev)
    s_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for a, b in s_evs:
        flush.zero_()
        a.record()
        step_e2e_serial()
        b.record()
    torch.cuda.synchronize(dev)
    e2e_serial_ms = float(sum(a.elapsed_time(b) for a, b in s_evs)) / e2e_steps

    # ---- secondary: chamfer fwd+bwd (configs[1]) ---------------------------------------------------
    ch = {k: v.to(dev) for k, v in make_chamfer_inputs(rank).items()}
    for k in ("x", "y", "xn", "yn", "xc", "yc"):
        ch[k].requires_grad_(True)

    def chamfer_step():
        for k in ("x", "y", "xn", "yn", "xc", "yc"):
            ch[k].grad = None
        loss, lf = chamfer_distance(ch["x"], ch["y"], x_lengths=ch["xl"], y_lengths=ch["yl"],
                                    x_features={"normals": ch["xn"], "colors": ch["xc"]},
                                    y_features={"normals": ch["yn"], "colors": ch["yc"]},
                                    feature_names=["normals", "colors"])
        (loss + lf["normals"] + lf["colors"]).backward()
        return loss

    for _ in range(3):
        chamfer_step()
    torch.cuda.synchronize(dev)
    c_steps = max(3, min(steps, 20))
    c_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(c_steps)]
    barrier()
    for a, b in c_evs:
        flush.zero_()
        a.record()
        chamfer_step()
        b.record()
    torch.cuda.synchronize(dev)
    barrier()
    c_ms = float(sum(a.elapsed_time(b) for a, b in c_evs))
    if dist is not None:
        t = torch.tensor([c_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c_ms = float(t.item())
    chamfer_pairs = 32 * world * c_steps / (c_ms * 1e-3)

    sms_early = torch.cuda.get_device_properties(dev).multi_processor_count
    # ---- secondary: high-D feature KNN on the tensor cores (configs[4]: D=128, K=16, B=16, P=32768;
    #      clouds shard 16/world per rank) ----------------------------------------------------------
    highdim = None
    if not args.no_highdim:
        Bh, Ph, Dh = max(1, 16 // world), 32768, 128
        gh = torch.Generator().manual_seed(4 + rank)
        xh = torch.randn(Bh, Ph, Dh, generator=gh).to(dev)
        lh = torch.full((Bh,), Ph, dtype=torch.int64, device=dev)
        for _ in range(3):
            _C.knn_points_idx(xh, xh, lh, lh, 2, K_NN, -1)
        torch.cuda.synchronize(dev)
        h_steps = 5
        h_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(h_steps)]
        lib.pops_profile_reset()
        lib.pops_profile_enable(1)
        barrier()
        for a, b in h_evs:
            a.record()
            _C.knn_points_idx(xh, xh, lh, lh, 2, K_NN, -1)
            b.record()
        torch.cuda.synchronize(dev)
        barrier()
        lib.pops_profile_enable(0)
        h_ms = float(sum(a.elapsed_time(b) for a, b in h_evs))
        if dist is not None:
            t = torch.tensor([h_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            h_ms = float(t.item())
        tc_ms, _ = kernel_ms(b"knn_tc_scan")
        rr_ms, _ = kernel_ms(b"knn_tc_rerank")
        ex_ms, _ = kernel_ms(b"knn_exact_rows")
        lib.pops_profile_reset()
        peaks_h, _ = load_peaks()
        tf32_peak = float(peaks_h.get("bf16_tflops", 1590.0)) / 2.0
        gemm_flop = 2.0 * Dh * Bh * Ph * Ph
        highdim = {
            "metric": "knn_queries_per_sec", "value": Bh * Ph * world * h_steps / (h_ms * 1e-3), "unit": UNIT,
            "ms_per_step": h_ms / h_steps,
            "workload": f"knn_points self-KNN B={Bh * world} ({Bh}/rank) P={Ph} D={Dh} K={K_NN} fp32 randn (configs[4]); "
                        "inputs 268 MB > L2, no flush needed",
            "kernels_ms": {"knn_tc_scan": tc_ms, "knn_tc_rerank": rr_ms, "knn_exact_rows": ex_ms},
            "roofline": {"kernel": "knn_tc_scan_kernel (tcgen05 kind::tf32)", "bound": "tensor",
                         "achieved": gemm_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None,
                         "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": gemm_flop / (tc_ms * 1e-3) / 1e12 / tf32_peak if tc_ms > 0 else None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (tf32 runs at half the bf16 rate; "
                                        "no measured tf32 entry)",
                         "algorithmic_flop_per_launch": gemm_flop},
        }
        del xh

    # ---- secondary: FPS (configs[2]: K=1024 from P=65536, B=64 sharded over 8 GPUs -> 64/8 per
    #      rank at world<=8) and ball query + gather (configs[3]: K=32, r=0.1, B=128, P=16384) ------
    others = {}
    if not args.no_highdim:
        from pytorch3d_pointops_b200.functions import ball_query, sample_farthest_points

        peaks_o, _ = load_peaks()
        Bf = 8  # configs[2]: 64 clouds over 8 GPUs = 8 per rank
        gf = torch.Generator().manual_seed(2 + rank)
        pf = torch.rand(Bf, 65536, 3, generator=gf).to(dev)
        for _ in range(2):
            sample_farthest_points(pf, K=1024)
        torch.cuda.synchronize(dev)
        f_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        barrier()
        for a, b in f_evs:
            a.record()
            sample_farthest_points(pf, K=1024)
            b.record()
        torch.cuda.synchronize(dev)
        f_ms = float(sum(a.elapsed_time(b) for a, b in f_evs)) / len(f_evs)
        f_bytes = Bf * 1023 * 65536 * (4 * 3 + 8)  # re-read points + r/w min-dist per iteration
        others["secondary_fps"] = {
            "metric": "fps_samples_per_sec", "value": Bf * 1024 * world / (f_ms * 1e-3), "unit": "samples/s",
            "ms_per_step": f_ms, "us_per_iteration": f_ms * 1e3 / 1023,
            "workload": f"sample_farthest_points K=1024 from P=65536, {Bf} clouds per rank (configs[2]: 64 clouds over 8 GPUs)",
            "roofline": {"kernel": "fps_d3_kernel (cluster per cloud, points in registers)", "bound": "hbm",
                         "achieved": f_bytes / (f_ms * 1e-3) / 1e9, "peak": float(peaks_o.get("hbm_gbs", 6650.0)),
                         "unit": "GB/s", "frac": f_bytes / (f_ms * 1e-3) / 1e9 / float(peaks_o.get("hbm_gbs", 6650.0)),
                         "note": "ALGORITHMIC bytes of the streaming formulation; the kernel keeps the cloud "
                                 "on chip and reads it from HBM once, so this is a latency-bound kernel"}}
        del pf
        Bb = max(1, 128 // world)
        gb = torch.Generator().manual_seed(3 + rank)
        pb = torch.rand(Bb, 16384, 3, generator=gb).to(dev)
        for _ in range(2):
            rb = ball_query(pb, pb, K=32, radius=0.1)
        torch.cuda.synchronize(dev)
        b_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        lib.pops_profile_reset()
        lib.pops_profile_enable(1)
        barrier()
        for a, b in b_evs:
            a.record()
            rb = ball_query(pb, pb, K=32, radius=0.1)
            b.record()
        torch.cuda.synchronize(dev)
        lib.pops_profile_enable(0)
        b_ms = float(sum(a.elapsed_time(b) for a, b in b_evs)) / len(b_evs)
        bq_ms, _ = kernel_ms(b"ball_query")
        lib.pops_profile_reset()
        last = rb.idx[..., -1]
        scanned = torch.where(last >= 0, last + 1, torch.full_like(last, 16384)).sum().item()
        others["secondary_ball_query"] = {
            "metric": "ball_query_queries_per_sec", "value": Bb * 16384 * world / (b_ms * 1e-3), "unit": UNIT,
            "ms_per_step": b_ms, "kernel_ms": bq_ms,
            "workload": f"ball_query K=32 r=0.1 return_nn=True (masked gather), B={Bb * world} ({Bb}/rank) P=16384 (configs[3])",
            "roofline": {"kernel": "ball_query_d3_kernel", "bound": "fp32",
                         "achieved": 9.0 * scanned / (bq_ms * 1e-3) / 1e12 if bq_ms > 0 else None,
                         "peak": sms_early * 128 * 2 * float(peaks_o.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12,
                         "unit": "TFLOP/s",
                         "note": "3*D flop per point the reference's sequential scan visits (idx[q,K-1]+1, or "
                                 "lengths2 when the ball holds fewer than K points), SURVEY 8(d)"}}
        rl = others["secondary_ball_query"]["roofline"]
        rl["frac"] = rl["achieved"] / rl["peak"] if rl["achieved"] else None
        del pb, rb

    # ---- roofline of the dominant kernel (knn_scan) -----------------------------------------------
    peaks, peak_src = load_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_theory = sms * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    fp32_probe = float(lib.pops_fp32_peak_probe(20000, torch.cuda.current_stream(dev).cuda_stream))
    alg_flops = 3.0 * D * pairs_per_step  # D sub + D mul + D add per pair
    achieved = alg_flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None
    roofline = {
        "kernel": "knn_prune_kernel<Q=1,KT=16> (Hilbert-ordered blocks, exact box pruning)", "bound": "fp32",
        "achieved": achieved,
        "peak": fp32_theory, "unit": "TFLOP/s", "frac": (achieved / fp32_theory) if achieved else None,
        "traffic": ncu_traffic_bytes("r01_knn_prune_ncu_full.txt"),
        "traffic_note": "DRAM bytes per launch from profiles/r01_knn_prune_ncu_full.txt (same shape); the "
                        "algorithmic minimum is 6.3 MB of points in + 100.7 MB of (idx, dists) out -- part of the output is still "
                        "in the 126 MB L2 when the kernel ends, so the capture can read below that",
        "note": "achieved = ALGORITHMIC flop (3*D per (query, point) pair of the brute-force definition, SURVEY 8d) "
                "/ kernel time; the kernel proves most blocks irrelevant and skips them, so frac can exceed "
                "what any brute-force scan reaches -- see executed_pair_fraction and bruteforce",
        "executed_pair_fraction": executed_fraction,
        "executed_tflops": (achieved * executed_fraction) if achieved else None,
        "bruteforce": {"what": "same kernel, knn_prune=0 (every block visited in the same order), knn_q=4 (4 queries per thread)",
                       "kernel_ms": brute_ms,
                       "achieved": alg_flops / (brute_ms * 1e-3) / 1e12 if brute_ms > 0 else None,
                       "frac": alg_flops / (brute_ms * 1e-3) / 1e12 / fp32_theory if brute_ms > 0 else None},
        "peak_source": f"SMs({sms}) x 128 FMA lanes x 2 x sm_max_mhz from {peak_src} (no FP32 entry there)",
        "peak_measured_ffma_probe": fp32_probe,
        "frac_of_probe": (achieved / fp32_probe) if (achieved and fp32_probe > 0) else None,
        "kernel_ms": scan_ms, "kernel_launches_timed": int(n_l.value),
        "algorithmic_flop_per_launch": alg_flops,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "384 MiB buffer written between timed steps (L2 flush)",
                   "pair_distance_evals_per_sec": pairs_per_step * world * steps / (total_ms * 1e-3),
                   "sharding": "by cloud: every rank owns its own batch, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / e2e_steps,
                "what": "pytorch3d_pointops_b200.host.HostKnn: pinned host clouds -> H2D -> knn_points_idx -> D2H of "
                        "dists+idx into pinned host, 8 slices of clouds: the first on its own, the rest behind one H2D and one pre-pass, each searched while the previous slice's results travel back (3 streams), replayed as one CUDA graph",
                "serial_ms_per_step": e2e_serial_ms,
                "serial_what": "p.to(device) -> knn_points -> copy_ of dists+idx to pinned host on one stream"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "secondary": {"metric": "chamfer_pairs_per_sec", "value": chamfer_pairs, "unit": "cloud-pairs/s",
                      "ms_per_step": c_ms / c_steps,
                      "workload": "chamfer_distance fwd+bwd B=32 P<=8192 ragged, normals+colors (configs[1])"},
        "wall_s_timed_region": wall1 - wall0,
    }
    if highdim is not None:
        line["secondary_highdim"] = highdim
    line.update(others)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, kind, sample, wall = cpu_reference_rate(p_host, len_host, budget_s=12.0, workers=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                                "seconds": wall}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-highdim", action="store_true", help="skip the D=128 tensor-core secondary measurement")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
