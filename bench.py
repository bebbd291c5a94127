#!/usr/bin/env python
"""bench.py -- headline benchmark of the nearest-neighbour hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference_cuda] [--sweep]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "KNN queries/sec (BxP1, K=16, D=3) and chamfer pairs/sec"):
  knn_T    : self-KNN, B=32 clouds x P=16384 points, D=3, K=16, fp32, uniform rand (seed 0+rank)
  chamfer  : chamfer_distance fwd+bwd, B=32, P<=8192 ragged, normals+colors ("secondary")
A step = one pass of the hot path over one batch.  Weak scaling: every rank owns its own batch
(clouds shard by batch index; no data-path collective), value = queries of all ranks / max time.

Arms
  ours            this repo's CUDA path (the JSON line the driver reads)
  reference       the reference's own CPU implementation (oracle/_ref) on all host cores
  reference_cuda  the UNMODIFIED reference CUDA extension (baseline/_ref, built for sm_100 from
                  /root/reference) on the same GPU: the secondary "GPU bar"; the ours arm runs it in
                  a subprocess at N=1 and reports `vs_reference_cuda` per op

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
# one dict for every arm, so that the driver's same_config check compares like with like
CONFIG = {
    "workload": WORKLOAD,
    "l2": "GPU arms: a 384 MiB buffer is written between timed steps (L2 flush); CPU arm: n/a",
    "sharding": "by cloud: every rank owns its own batch, no data-path collective",
}
HBM_FALLBACK = 6650.0


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

    def __init__(self, index: int, period: float = 0.002):
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
                "samples": len(s), "source": "nvml, sampled during the timed region"}

    def _smi_fallback(self):
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


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs next to its GPU, so that the pinned host buffers it allocates are
    NUMA-local (first touch) -- with 8 ranks copying 100 MB per step concurrently, remote pinned memory
    shares one inter-socket link.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return sorted(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------------
# inputs (identical for every arm)
# --------------------------------------------------------------------------------------------
def make_knn_inputs(rank: int, ragged: bool = False):
    g = torch.Generator().manual_seed(0 + rank)
    p = torch.rand(B, P, D, generator=g)
    lengths = torch.full((B,), P, dtype=torch.int64)
    if ragged:
        lengths = torch.randint(8192, P + 1, (B,), generator=g)
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


def make_fps_inputs(rank: int, clouds: int = 8):
    g = torch.Generator().manual_seed(2 + rank)
    return torch.rand(clouds, 65536, 3, generator=g)


def make_ball_inputs(rank: int, clouds: int):
    g = torch.Generator().manual_seed(3 + rank)
    return torch.rand(clouds, 16384, 3, generator=g)


def ncu_dram_bytes(summary_name: str, kernel_substr: str = None):
    """dram__bytes_read.sum + dram__bytes_write.sum of a committed `ncu --set full` summary
    (profiles/<summary_name>), for the first kernel whose name (or the "# shape" note under it) contains
    kernel_substr; or None."""
    path = os.path.join(REPO, "profiles", summary_name)
    if not os.path.isfile(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen, active = 0.0, 0, kernel_substr is None
    with open(path) as fh:
        for ln in fh:
            ln = ln.strip()
            if ln.startswith("kernel:"):
                if seen == 2:
                    break
                active = kernel_substr is None or kernel_substr in ln
                total, seen = 0.0, 0
            elif ln.startswith("#") and kernel_substr and kernel_substr in ln:
                active = True
            elif active and ln.startswith(("dram__bytes_read.sum =", "dram__bytes_write.sum =")):
                val, u = ln.split("=")[1].split()[:2]
                total += float(val) * unit.get(u, 1.0)
                seen += 1
    return total if seen == 2 else None


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh), "MEASURED_PEAKS.json (measured)"
    return {"hbm_gbs": HBM_FALLBACK, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


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
    n, dt = _ref_worker((0, 0, 128, use_ref))  # calibrate on one core: 128 queries of cloud 0
    rate1 = n / max(dt, 1e-6)
    per_worker = int(max(64, rate1 * budget_s))          # queries one core finishes in the budget
    jobs = []
    for w in range(workers):  # worker w walks clouds w, w+workers, ...
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
        "data": "synthetic", "config": dict(CONFIG),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# shared timing helpers (GPU arms)
# --------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, dev, dist=None, local_rank=0):
        self.dev, self.dist, self.local_rank = dev, dist, local_rank
        self.flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier(device_ids=[self.local_rank])
        torch.cuda.synchronize(self.dev)

    def allmax(self, v: float) -> float:
        if self.dist is None:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, fn, steps: int, warmup: int = 3, flush: bool = True, reduce: bool = True, lead: bool = False):
        """ms per step of fn: `warmup` untimed calls, then `steps` calls bracketed by CUDA events on the
        current stream with the L2 flushed before each, barrier + synchronize on both sides, max over
        ranks of the total.  lead: park the GPU for ~0.2 ms in front of every call, so that the host has
        queued the call's launches before the device reaches them -- for kernels timed by the library's
        event hooks (pops_profile_*), whose begin event would otherwise also cover the host's launch gap
        when fn ends with a host sync (knn_gather's index check does)."""
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(self.dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in evs:
            if flush:
                self.flush.zero_()
            if lead:
                torch.cuda._sleep(400_000)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize(self.dev)
        self.barrier()
        total = float(sum(a.elapsed_time(b) for a, b in evs))
        return (self.allmax(total) if reduce else total) / steps


def hbm_line(kernel, ms, alg_bytes, dram_bytes, peak, peak_src, traffic=None, note=None):
    """roofline object of an HBM-bound kernel: `achieved` on the ALGORITHMIC bytes of SURVEY.md 8(d),
    `frac_dram` on the bytes that must cross HBM at least once (the gathered rows of a cloud are
    re-read from L2, not from HBM)."""
    out = {"kernel": kernel, "bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else None,
           "peak": peak, "unit": "GB/s", "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
           "algorithmic_bytes_per_launch": alg_bytes, "compulsory_dram_bytes_per_launch": dram_bytes,
           "frac_dram": dram_bytes / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
           "traffic": traffic, "peak_source": peak_src, "kernel_ms": ms}
    if note:
        out["note"] = note
    return out


# --------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    cpus = bind_to_gpu_numa_node(local_rank)
    from pytorch3d_pointops_b200 import _C, _lib
    from pytorch3d_pointops_b200.functions import knn_points
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

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
    tm = Timer(dev, dist, local_rank)
    barrier, flush = tm.barrier, tm.flush
    peaks, peak_src = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK))
    import ctypes

    def kernel_ms(name):
        n_, ms_ = ctypes.c_int64(0), ctypes.c_double(0.0)
        lib.pops_profile_read(name, ctypes.byref(n_), ctypes.byref(ms_))
        return ms_.value / max(1, n_.value), int(n_.value)

    steps, warm = max(1, args.steps), max(3, args.warmup)
    p_host, len_host = make_knn_inputs(rank)
    p_pin = p_host.pin_memory()
    p_dev = p_host.to(dev)
    len_dev = len_host.to(dev)
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
    total_ms = tm.allmax(float(sum(a.elapsed_time(b) for a, b in evs)))
    value = queries_per_step * world * steps / (total_ms * 1e-3)
    scan_ms, scan_launches = kernel_ms(b"knn_scan")
    lib.pops_profile_reset()

    # ---- what the pruned search executes: block counters of one untimed step, and the same kernel
    #      with pruning switched off (every block visited, same order) as the brute-force reference
    lib.pops_set_option(b"knn_stats", 1)
    stats = (ctypes.c_ulonglong * 8)()
    lib.pops_knn_debug_stats(stats)
    step_resident()
    lib.pops_knn_debug_stats(stats)
    lib.pops_set_option(b"knn_stats", 0)
    warps = max(1, int(stats[5]))
    # per query: the 16-point runs its warp scans (filter form), the 3 seed blocks (64 points each), and the
    # buffered groups of 4 points that get the exact distance (a warp = 32 queries)
    executed_fraction = ((int(stats[6]) / warps) * 16.0 + 3 * 64.0 + (int(stats[3]) / warps) * 4.0 / 32.0) / P
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

    # ---- ragged T shape (lengths = randint(8192, 16385)) ------------------------------------
    pr_host, lr_host = make_knn_inputs(rank, ragged=True)
    pr_dev, lr_dev = pr_host.to(dev), lr_host.to(dev)
    r_steps = max(3, min(steps, 20))
    lib.pops_profile_enable(1)
    r_ms = tm.run(lambda: _C.knn_points_idx(pr_dev, pr_dev, lr_dev, lr_dev, 2, K_NN, -1), r_steps)
    lib.pops_profile_enable(0)
    r_scan_ms, _ = kernel_ms(b"knn_scan")
    lib.pops_profile_reset()
    ragged_queries = int(lr_host.sum())
    secondary_ragged = {
        "metric": METRIC, "value": ragged_queries * world / (r_ms * 1e-3), "unit": UNIT, "ms_per_step": r_ms,
        "kernel_ms": r_scan_ms, "queries_per_step_per_rank": ragged_queries,
        "pair_distance_evals_per_sec": int((lr_host * lr_host).sum()) * world / (r_ms * 1e-3),
        "workload": f"knn_points self-KNN B={B} P<={P} ragged (lengths = randint(8192, {P + 1}), seed 0+rank) D={D} K={K_NN}",
    }
    del pr_dev

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
    e2e_steps = max(3, min(steps, 20))
    e2e_ms = tm.run(lambda: host_knn(p_pin, None, len_pin), e2e_steps, warmup=2)
    e2e_value = queries_per_step * world / (e2e_ms * 1e-3)
    h2d = p_pin.numel() * 4 + len_pin.numel() * 8
    d2h = out_d_pin.numel() * 4 + out_i_pin.numel() * 8
    # the plain call sequence a user of the reference API writes (no overlap), for comparison
    e2e_serial_ms = tm.run(step_e2e_serial, e2e_steps, warmup=2, reduce=False)
    # the box's own copy floors: every rank moves the SAME bytes with plain cudaMemcpyAsync (one per
    # buffer) into / out of pinned memory at the same time, nothing else running
    dd = torch.empty((B, P, K_NN), dtype=torch.float32, device=dev)
    di = torch.empty((B, P, K_NN), dtype=torch.int64, device=dev)

    def d2h_only():
        out_d_pin.copy_(dd, non_blocking=True)
        out_i_pin.copy_(di, non_blocking=True)

    def h2d_only():
        p_dev.copy_(p_pin, non_blocking=True)
        len_dev.copy_(len_pin, non_blocking=True)

    d2h_floor_ms = tm.run(d2h_only, e2e_steps, warmup=2, flush=False)
    h2d_floor_ms = tm.run(h2d_only, e2e_steps, warmup=2, flush=False)
    # additive variant: 32-bit indices on the wire (a third less D2H); the contract line stays int64
    host_knn32 = HostKnn(B, P, P, D, K_NN, dev, slices=8, idx_dtype=torch.int32)
    e2e32_ms = tm.run(lambda: host_knn32(p_pin, None, len_pin), e2e_steps, warmup=2)
    del dd, di, host_knn32

    # ---- secondary: chamfer fwd+bwd (configs[1]) ---------------------------------------------------
    ch = {k: v.to(dev) for k, v in make_chamfer_inputs(rank).items()}
    grads = ("x", "y", "xn", "yn", "xc", "yc")
    for k in grads:
        ch[k].requires_grad_(True)

    def chamfer_call(fn, c, **extra):
        loss, lf = fn(c["x"], c["y"], x_lengths=c["xl"], y_lengths=c["yl"],
                      x_features={"normals": c["xn"], "colors": c["xc"]},
                      y_features={"normals": c["yn"], "colors": c["yc"]},
                      feature_names=["normals", "colors"], **extra)
        return loss, lf

    def chamfer_step():
        for k in grads:
            ch[k].grad = None
        loss, lf = chamfer_call(chamfer_distance, ch)
        (loss + lf["normals"] + lf["colors"]).backward()
        return loss

    c_steps = max(3, min(steps, 20))
    c_launch0 = _lib.launch_count()
    c_ms = tm.run(chamfer_step, c_steps)
    c_launches = (_lib.launch_count() - c_launch0) // (c_steps + 3)

    def chamfer_step_fresh_lengths():  # a caller that builds new lengths tensors every step pays the validation sync
        for k in grads:
            ch[k].grad = None
        c2 = dict(ch)
        c2["xl"], c2["yl"] = ch["xl"].clone(), ch["yl"].clone()
        loss, lf = chamfer_call(chamfer_distance, c2)
        (loss + lf["normals"] + lf["colors"]).backward()

    c_fresh_ms = tm.run(chamfer_step_fresh_lengths, c_steps, reduce=False)
    secondary = {"metric": "chamfer_pairs_per_sec", "value": 32 * world / (c_ms * 1e-3), "unit": "cloud-pairs/s",
                 "ms_per_step": c_ms, "launches_per_step": int(c_launches),
                 "ms_per_step_fresh_lengths": c_fresh_ms,
                 "workload": "chamfer_distance fwd+bwd B=32 P<=8192 ragged, normals+colors (configs[1]); per-rank loss, no collective",
                 "note": "steady state reuses the lengths tensors (their max() was validated once); "
                         "ms_per_step_fresh_lengths passes new lengths tensors every step (one host sync each)"}
    if dist is not None:
        # the one collective of the path: all-reduce of (1 + #features) scalars so that every rank holds
        # the GLOBAL batch-mean loss (distributed.chamfer_distance_sharded), inside the timed region
        from pytorch3d_pointops_b200.distributed import chamfer_distance_sharded

        def chamfer_step_sharded():
            for k in grads:
                ch[k].grad = None
            loss, lf = chamfer_call(chamfer_distance_sharded, ch, n_clouds_global=32 * world)
            (loss + lf["normals"] + lf["colors"]).backward()
            return loss, lf

        cs_ms = tm.run(lambda: chamfer_step_sharded(), c_steps)
        ar_buf = torch.zeros(4, device=dev)
        ar_ms = tm.run(lambda: dist.all_reduce(ar_buf), 20, flush=False)
        # the sharded loss must equal the single-GPU loss on the gathered batch (every rank can rebuild
        # all shards from the seeds)
        loss_s, lf_s = chamfer_step_sharded()
        full = [make_chamfer_inputs(r) for r in range(world)]
        cat = {k: torch.cat([f[k] for f in full], 0).to(dev) for k in full[0]}
        with torch.no_grad():
            loss_f, lf_f = chamfer_call(chamfer_distance, cat)
        ok = bool(torch.allclose(loss_s.detach(), loss_f, rtol=1e-5) and
                  all(torch.allclose(lf_s[k].detach(), lf_f[k], rtol=1e-5) for k in lf_f))
        assert ok, ("sharded chamfer loss differs from the single-GPU loss on the gathered batch",
                    float(loss_s), float(loss_f))
        secondary["sharded"] = {
            "value": 32 * world / (cs_ms * 1e-3), "unit": "cloud-pairs/s", "ms_per_step": cs_ms,
            "collective": f"NCCL all-reduce (sum) of 4 fp32 scalars over {world} ranks, inside the timed region",
            "collective_us": ar_ms * 1e3, "added_us_vs_local_step": (cs_ms - c_ms) * 1e3,
            "loss_equals_single_gpu_loss_on_gathered_batch_rtol_1e-5": ok,
            "loss": float(loss_s.detach()), "loss_gathered": float(loss_f)}
        del cat, full

    # ---- secondary: the HBM-bound rows (SURVEY 8d byte formulas) -------------------------------------
    hbm = {}
    if not args.no_hbm:
        from pytorch3d_pointops_b200.functions import ball_query, knn_gather
        from pytorch3d_pointops_b200.functions.packed_to_padded import packed_to_padded, padded_to_packed

        h_steps = 10

        def hooked(fn, name, steps=h_steps):
            """average device time of the library kernel `name` over `steps` calls of fn, by the library's event
            hooks.  Warm-up runs with the hooks OFF (a call that ends with a host sync leaves the device idle,
            and the next call's begin event would cover the host's launch gap), timed calls get a device lead."""
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            lib.pops_profile_reset()
            lib.pops_profile_enable(1)
            tm.run(fn, steps, warmup=0, reduce=False, lead=True)
            lib.pops_profile_enable(0)
            ms, _ = kernel_ms(name)
            lib.pops_profile_reset()
            return ms

        # knn_gather on the T shape's KNN indices (U = 3: xyz; U = 16: a feature row)
        idx_T, _ = step_resident()
        rows = B * P * K_NN
        g3_ms = hooked(lambda: knn_gather(p_dev, idx_T, len_dev), b"gather")
        feat16 = torch.rand(B, P, 16, device=dev)
        g16_ms = hooked(lambda: knn_gather(feat16, idx_T, len_dev), b"gather")
        hbm["secondary_gather"] = {
            "metric": "gathered_rows_per_sec", "value": rows * world / (g3_ms * 1e-3), "unit": "rows/s",
            "workload": f"knn_gather of the T shape's KNN indices: x ({B},{P},U) f32, idx ({B},{P},{K_NN}) i64",
            "roofline": hbm_line("gather_rows3_smem_kernel<KNN> (U=3, the cloud staged in shared memory)", g3_ms,
                                 rows * (8 + 12 + 12), rows * (8 + 12) + B * P * 12,
                                 hbm_peak, peak_src, ncu_dram_bytes("r02_hbm_kernels_ncu_full.txt", "T shape: knn_gather")),
            "u16": hbm_line("gather_kernel<KNN,V4> U=16", g16_ms, rows * (8 + 64 + 64), rows * (8 + 64) + B * P * 64,
                            hbm_peak, peak_src)}
        del feat16
        # KNN backward on the T shape
        gd = torch.rand(B, P, K_NN, device=dev)
        kb_ms = hooked(lambda: _C.knn_points_backward(p_dev, p_dev, len_dev, len_dev, idx_T, 2, gd), b"knn_backward")
        hbm["secondary_knn_backward"] = {
            "metric": "knn_backward_entries_per_sec", "value": rows * world / (kb_ms * 1e-3), "unit": "(query,neighbour) pairs/s",
            "workload": f"_C.knn_points_backward on the T shape: idx/grad_dists ({B},{P},{K_NN}), D=3, norm 2",
            "roofline": hbm_line("knn_backward_rows_kernel<2,3> + compact", kb_ms,
                                 rows * 12 + 2 * rows * 3 * 4 + 2 * B * P * 3 * 4, rows * 12 + 3 * B * P * 12,
                                 hbm_peak, peak_src, ncu_dram_bytes("r02_hbm_kernels_ncu_full.txt", "knn_backward_rows"),
                                 note="bounded by L2 reduction throughput (one 16-byte red per (query, neighbour)), not by HBM")}
        del gd, idx_T
        # packed <-> padded: 64 ragged clouds of 32768..65536 points
        gpk = torch.Generator().manual_seed(5 + rank)
        lens = torch.randint(32768, 65537, (64,), generator=gpk)
        first = (torch.cumsum(lens, 0) - lens).to(dev)
        F_rows, mx = int(lens.sum()), int(lens.max())
        pk = {}
        for Dp in (3, 16):
            packed = torch.rand(F_rows, Dp, device=dev)
            a_ms = hooked(lambda: packed_to_padded(packed, first, mx), b"packed_to_padded")
            padded = packed_to_padded(packed, first, mx)
            b_ms = hooked(lambda: padded_to_packed(padded, first, F_rows), b"padded_to_packed")
            by_a = F_rows * Dp * 4 + 64 * mx * Dp * 4
            by_b = 2 * F_rows * Dp * 4
            pk[f"D{Dp}"] = {"packed_to_padded": hbm_line("packed_to_padded_kernel", a_ms, by_a, by_a, hbm_peak, peak_src,
                                                         ncu_dram_bytes("r02_hbm_kernels_ncu_full.txt", "packed_to_padded") if Dp == 3 else None),
                            "padded_to_packed": hbm_line("padded_to_packed_kernel", b_ms, by_b, by_b, hbm_peak, peak_src,
                                                         ncu_dram_bytes("r02_hbm_kernels_ncu_full.txt", "padded_to_packed") if Dp == 3 else None)}
            del packed, padded
        hbm["secondary_pack"] = {
            "metric": "packed_to_padded_GBps", "value": pk["D3"]["packed_to_padded"]["achieved"], "unit": "GB/s",
            "workload": f"packed_to_padded / padded_to_packed, 64 ragged clouds of 32768..65536 rows (F={F_rows}), D=3 and D=16",
            "roofline": pk["D3"]["packed_to_padded"], "all": pk}

    # ---- secondary: high-D feature KNN on the tensor cores (configs[4]: D=128, K=16, B=16, P=32768;
    #      clouds shard 16/world per rank) ----------------------------------------------------------
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_theory = sms * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    highdim = None
    others = {}
    if not args.no_highdim:
        Bh, Ph, Dh = max(1, 16 // world), 32768, 128
        gh = torch.Generator().manual_seed(4 + rank)
        xh = torch.randn(Bh, Ph, Dh, generator=gh).to(dev)
        lh = torch.full((Bh,), Ph, dtype=torch.int64, device=dev)
        lib.pops_profile_reset()
        lib.pops_profile_enable(1)
        h_ms = tm.run(lambda: _C.knn_points_idx(xh, xh, lh, lh, 2, K_NN, -1), 5, flush=False)
        lib.pops_profile_enable(0)
        tc_ms, _ = kernel_ms(b"knn_tc_scan")
        rr_ms, _ = kernel_ms(b"knn_tc_rerank")
        ex_ms, _ = kernel_ms(b"knn_exact_rows")
        lib.pops_profile_reset()
        tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        gemm_flop = 2.0 * Dh * Bh * Ph * Ph
        highdim = {
            "metric": METRIC, "value": Bh * Ph * world / (h_ms * 1e-3), "unit": UNIT, "ms_per_step": h_ms,
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

        # ---- FPS (configs[2]: K=1024 from P=65536, B=64 sharded over 8 GPUs -> 8 per rank) and
        #      ball query + gather (configs[3]: K=32, r=0.1, B=128, P=16384) ---------------------------
        from pytorch3d_pointops_b200.functions import ball_query, sample_farthest_points

        Bf = 8
        pf = make_fps_inputs(rank, Bf).to(dev)
        f_ms = tm.run(lambda: sample_farthest_points(pf, K=1024), 5, warmup=2, flush=False, reduce=False)
        f_bytes = Bf * 1023 * 65536 * (4 * 3 + 8)  # re-read points + r/w min-dist per iteration
        others["secondary_fps"] = {
            "metric": "fps_us_per_iteration", "value": f_ms * 1e3 / 1023, "unit": "us/iteration", "higher_is_better": False,
            "samples_per_sec": Bf * 1024 * world / (f_ms * 1e-3), "ms_per_step": f_ms,
            "workload": f"sample_farthest_points K=1024 from P=65536, {Bf} clouds per rank (configs[2]: 64 clouds over 8 GPUs)",
            "roofline": {"kernel": "fps_d3_kernel (cluster per cloud, points in registers)", "bound": "latency",
                         "note": "on-chip kernel: the cloud is read from HBM once (6.3 MB); the dependent arg-max chain "
                                 "bounds it, so the headline is us/iteration.  streaming_equivalent = the bytes the "
                                 "reference's formulation re-reads per iteration (SURVEY 8d) / time -- NOT HBM traffic",
                         "streaming_equivalent_GBps": f_bytes / (f_ms * 1e-3) / 1e9,
                         "streaming_equivalent_frac_of_hbm": f_bytes / (f_ms * 1e-3) / 1e9 / hbm_peak}}
        del pf
        Bb = max(1, 128 // world)
        pb = make_ball_inputs(rank, Bb).to(dev)

        def bq_timed():
            lib.pops_profile_reset()
            lib.pops_profile_enable(1)
            ms = tm.run(lambda: ball_query(pb, pb, K=32, radius=0.1), 5, warmup=2, flush=False, reduce=False)
            lib.pops_profile_enable(0)
            kms, _ = kernel_ms(b"ball_query")
            lib.pops_profile_reset()
            return ms, kms

        b_ms, bq_ms = bq_timed()              # default: per-cloud choice (this shape: the Hilbert-ordered search)
        lib.pops_set_option(b"bq_spatial", 0)  # the index-order scan alone: the kernel the FP32 fraction belongs to
        s_ms, sq_ms = bq_timed()
        lib.pops_set_option(b"bq_spatial", -1)
        rb = ball_query(pb, pb, K=32, radius=0.1, return_nn=False)
        last = rb.idx[..., -1]
        scanned = torch.where(last >= 0, last + 1, torch.full_like(last, 16384)).sum().item()
        ref_flop = 9.0 * scanned
        others["secondary_ball_query"] = {
            "metric": "ball_query_queries_per_sec", "value": Bb * 16384 * world / (b_ms * 1e-3), "unit": UNIT,
            "ms_per_step": b_ms, "kernel_ms": bq_ms,
            "workload": f"ball_query K=32 r=0.1 return_nn=True (masked gather), B={Bb * world} ({Bb}/rank) P=16384 (configs[3])",
            "kernels": "bq_decide_kernel (sampled hit counts pick the kernel per cloud) + bq_prune_kernel (Hilbert-ordered "
                       "blocks: all hits through the block boxes, K smallest indices by warp-wide bitonic sorts) + "
                       "ball_query_scan_kernel (clouds the ordered search did not take); kernel_ms covers the three, "
                       "ms_per_step adds the ordering pre-pass and the masked gather",
            "roofline": {"kernel": "ball_query_scan_kernel (bq_spatial=0: every cloud through the index-order scan)",
                         "bound": "fp32", "kernel_ms": sq_ms, "ms_per_step": s_ms,
                         "achieved": ref_flop / (sq_ms * 1e-3) / 1e12 if sq_ms > 0 else None,
                         "peak": fp32_theory, "unit": "TFLOP/s",
                         "note": "3*D flop per point the reference's sequential scan visits (idx[q,K-1]+1, or "
                                 "lengths2 when the ball holds fewer than K points), SURVEY 8(d); the scan kernel "
                                 "visits at least these (a CTA runs until its slowest query is complete)"},
            "algorithmic_tflops": ref_flop / (bq_ms * 1e-3) / 1e12 if bq_ms > 0 else None,
            "algorithmic_note": "the same reference-scan flop / the DEFAULT path's kernel time: how its time compares "
                                "with the index-order scan at a given FP32 rate.  Not a hardware fraction -- the "
                                "ordered search evaluates only the points of blocks within the radius",
            "speedup_vs_index_scan": s_ms / b_ms if b_ms > 0 else None}
        rl = others["secondary_ball_query"]["roofline"]
        rl["frac"] = rl["achieved"] / rl["peak"] if rl["achieved"] else None
        del pb, rb

    # ---- roofline of the dominant kernel (knn_scan) -----------------------------------------------
    fp32_probe = float(lib.pops_fp32_peak_probe(20000, torch.cuda.current_stream(dev).cuda_stream))
    alg_flops = 3.0 * D * pairs_per_step  # D sub + D mul + D add per pair
    alg_tflops = alg_flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None
    exe_tflops = alg_tflops * executed_fraction if alg_tflops else None
    roofline = {
        "kernel": "knn_prune_kernel<Q=1,KT=16> (Hilbert-ordered blocks, exact box pruning)", "bound": "fp32",
        "achieved": exe_tflops, "peak": fp32_theory, "unit": "TFLOP/s",
        "frac": (exe_tflops / fp32_theory) if exe_tflops else None,
        "what": "EXECUTED work: 3*D flop for every (query, point) pair the kernel actually evaluates "
                "(16-point runs scanned + 3 seed blocks + exact re-evaluations, from its own counters) / kernel time.  The kernel is "
                "latency / issue bound, not FP32-pipe bound: most of its instructions are selection, not distance arithmetic",
        "executed_pair_fraction": executed_fraction,
        "algorithmic_tflops": alg_tflops,
        "algorithmic_speedup": (alg_tflops / fp32_theory) if alg_tflops else None,
        "algorithmic_note": "ALGORITHMIC flop of the brute-force definition (3*D per pair of lengths1 x lengths2, "
                            "SURVEY 8d) / kernel time / FP32 peak: how the time compares with a brute-force scan "
                            "running AT the FP32 peak; above 1 because exact bounding-box pruning proves most "
                            "pairs irrelevant.  Not a hardware fraction",
        "traffic": ncu_dram_bytes("r02_knn_prune_ncu_full.txt") or ncu_dram_bytes("r01_knn_prune_ncu_full.txt"),
        "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (same shape); algorithmic minimum "
                        "6.3 MB of points in + 100.7 MB of (idx, dists) out, part of which is still in the 126 MB L2 "
                        "when the kernel ends",
        "bruteforce": {"what": "same kernel, knn_prune=0 (every block visited in the same order), knn_q=4 (4 queries per thread)",
                       "kernel_ms": brute_ms,
                       "achieved": alg_flops / (brute_ms * 1e-3) / 1e12 if brute_ms > 0 else None,
                       "frac": alg_flops / (brute_ms * 1e-3) / 1e12 / fp32_theory if brute_ms > 0 else None},
        "peak_source": f"SMs({sms}) x 128 FMA lanes x 2 x sm_max_mhz from {peak_src} (no FP32 entry there)",
        "peak_measured_ffma_probe": fp32_probe,
        "kernel_ms": scan_ms, "kernel_launches_timed": int(scan_launches),
        "kernel_share_of_step": scan_ms / (total_ms / steps) if total_ms > 0 else None,
        "algorithmic_flop_per_launch": alg_flops,
    }

    config = dict(CONFIG)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "pair_distance_evals_per_sec": pairs_per_step * world * steps / (total_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms,
                "what": "pytorch3d_pointops_b200.host.HostKnn: pinned host clouds -> H2D -> knn_points_idx -> D2H of "
                        "dists+idx (int64) into pinned host, 8 slices of clouds: the first on its own, the rest behind one H2D "
                        "and one pre-pass, each searched while the previous slice's results travel back (3 streams), "
                        "replayed as one CUDA graph",
                "serial_ms_per_step": e2e_serial_ms,
                "serial_what": "p.to(device) -> knn_points -> copy_ of dists+idx to pinned host on one stream",
                "d2h_floor_ms": d2h_floor_ms, "h2d_floor_ms": h2d_floor_ms,
                "floor_what": f"all {world} rank(s) copying the same {d2h} B (D2H: one cudaMemcpyAsync per buffer) / {h2d} B (H2D) "
                              "to / from pinned memory at the same time, nothing else running; max over ranks",
                "d2h_floor_over_e2e": d2h_floor_ms / e2e_ms if e2e_ms > 0 else None,
                "cpu_affinity": f"{len(cpus)} CPUs next to the GPU (NVML)" if cpus else "not set",
                "int32_idx": {"value": queries_per_step * world / (e2e32_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e32_ms,
                              "d2h_bytes_per_step": out_d_pin.numel() * 4 + out_i_pin.numel() * 4,
                              "what": "additive HostKnn(idx_dtype=torch.int32): indices narrowed on the device before the "
                                      "D2H; NOT the reference contract (int64) -- reported beside it, never instead"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "secondary": secondary,
        "secondary_ragged": secondary_ragged,
        "wall_s_timed_region": wall1 - wall0,
    }
    if highdim is not None:
        line["secondary_highdim"] = highdim
    line.update(hbm)
    line.update(others)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, kind, sample, wall = cpu_reference_rate(p_host, len_host, budget_s=12.0, workers=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                                "seconds": wall}
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        line["vs_reference_cuda"] = reference_cuda_ratios(line, args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# the GPU bar: the unmodified reference CUDA extension on the same box
# --------------------------------------------------------------------------------------------
REF_CUDA_DIR = os.path.join(REPO, "baseline", "_ref")


def reference_cuda_ratios(line, args):
    """Run `bench.py --impl reference_cuda` in a subprocess (its `pytorch3d_pointops` must not meet this
    repo's alias package of the same name) and divide: > 1 = this repo is faster."""
    if not os.path.isdir(os.path.join(REF_CUDA_DIR, "pytorch3d_pointops")):
        return {"unavailable": "baseline/_ref not present (pip install of /root/reference with FORCE_CUDA=1, see DESIGN.md)"}
    torch.cuda.empty_cache()
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference_cuda"],
                           capture_output=True, text=True, timeout=900)
        ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"reference_cuda arm failed: {e}"}
    if "unavailable" in ref:
        return ref
    ours = {"knn_T": line["ms_per_step"], "chamfer_C2": line["secondary"]["ms_per_step"]}
    if "secondary_fps" in line:
        ours["fps_C3_share"] = line["secondary_fps"]["ms_per_step"]
    if "secondary_ball_query" in line:
        ours["ball_query_C4"] = line["secondary_ball_query"]["ms_per_step"]
    out = {"what": "reference CUDA ms / this repo's ms on the same tensors, same GPU, same process order "
                   "(device-resident, CUDA events, median of 5 after 2 warm-ups); > 1 = this repo is faster",
           "reference_build": ref.get("build")}
    for k, v in ours.items():
        if k in ref["ms"]:
            out[k] = {"reference_cuda_ms": ref["ms"][k], "ours_ms": v, "speedup": ref["ms"][k] / v}
    return out


def run_reference_cuda(args):
    """The UNMODIFIED reference package from baseline/_ref (FORCE_CUDA=1 TORCH_CUDA_ARCH_LIST=10.0 build of
    /root/reference) through its own public API, on T, C2, the C3 per-rank share and C4."""
    pkg = os.path.join(REF_CUDA_DIR, "pytorch3d_pointops")
    if not os.path.isdir(pkg) or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference_cuda", "unavailable": "baseline/_ref or a CUDA device is missing"}), flush=True)
        return
    sys.path[:] = [REF_CUDA_DIR] + [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    try:
        import pytorch3d_pointops as ref
        from pytorch3d_pointops import _C as ref_C  # noqa: F401
        from pytorch3d_pointops.functions import ball_query, knn_points, sample_farthest_points
        from pytorch3d_pointops.functions.chamfer import chamfer_distance
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference_cuda", "unavailable": f"import failed: {e}"}), flush=True)
        return
    assert ref.__file__.startswith(REF_CUDA_DIR), ref.__file__
    dev = torch.device("cuda", 0)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def med(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return sorted(ms)[len(ms) // 2]

    out = {}
    if args.sweep:
        print(json.dumps({"impl": "reference_cuda", "sweep": sweep_table(knn_points, ball_query, sample_farthest_points, dev)}),
              flush=True)
        return
    p, L = make_knn_inputs(0)
    pd, Ld = p.to(dev), L.to(dev)
    out["knn_T"] = med(lambda: knn_points(pd, pd, Ld, Ld, K=K_NN))
    del pd
    ch = {k: v.to(dev) for k, v in make_chamfer_inputs(0).items()}
    grads = ("x", "y", "xn", "yn", "xc", "yc")
    for k in grads:
        ch[k].requires_grad_(True)

    def chamfer_step():
        for k in grads:
            ch[k].grad = None
        loss, lf = chamfer_distance(ch["x"], ch["y"], x_lengths=ch["xl"], y_lengths=ch["yl"],
                                    x_features={"normals": ch["xn"], "colors": ch["xc"]},
                                    y_features={"normals": ch["yn"], "colors": ch["yc"]},
                                    feature_names=["normals", "colors"])
        (loss + lf["normals"] + lf["colors"]).backward()

    out["chamfer_C2"] = med(chamfer_step)
    del ch
    pf = make_fps_inputs(0, 8).to(dev)
    out["fps_C3_share"] = med(lambda: sample_farthest_points(pf, K=1024), reps=3, warm=1)
    del pf
    pb = make_ball_inputs(0, 128).to(dev)
    out["ball_query_C4"] = med(lambda: ball_query(pb, pb, K=32, radius=0.1), reps=3, warm=1)
    print(json.dumps({"impl": "reference_cuda", "ms": out,
                      "build": "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref of a /tmp copy "
                               "of /root/reference with FORCE_CUDA=1 TORCH_CUDA_ARCH_LIST=10.0 (unmodified sources)"}),
          flush=True)


# --------------------------------------------------------------------------------------------
# --sweep: the size sweeps of the reference's timing harness (examples/cuda_vs_python_performance.py:
# KNN :122-124, ball query :171-178, FPS :223-224, batching :369-371) -- the small-P / small-B regime
# --------------------------------------------------------------------------------------------
def sweep_table(knn_points, ball_query, sample_farthest_points, dev):
    def avg_ms(fn, runs=10, warmup=3):  # the harness's own protocol: wall clock, synchronise per call
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(runs):
            fn()
            torch.cuda.synchronize()
        return (time.perf_counter() - t0) / runs * 1e3

    g = torch.Generator().manual_seed(42)
    table = {"knn_K16": {}, "ball_query_r0.5_K20": {}, "fps_10pct": {}, "batch_P500_K16": {}}
    for size in (100, 500, 1000, 2000, 10000, 32000):
        x = torch.randn(1, size, 3, generator=g).to(dev)
        table["knn_K16"][str(size)] = avg_ms(lambda: knn_points(x, x, K=16, return_nn=False))
    for size in (100, 500, 1000, 10000):
        x = torch.randn(1, size, 3, generator=g).to(dev)
        table["ball_query_r0.5_K20"][str(size)] = avg_ms(lambda: ball_query(x, x, K=20, radius=0.5, return_nn=False))
    for size in (500, 1000, 2000, 5000):
        x = torch.randn(1, size, 3, generator=g).to(dev)
        table["fps_10pct"][str(size)] = avg_ms(lambda: sample_farthest_points(x, K=int(size * 0.1), random_start_point=False))
    for bs in (1, 2, 4, 8, 16, 32):
        x = torch.randn(bs, 500, 3, generator=g).to(dev)
        L = torch.full((bs,), 500, dtype=torch.int64, device=dev)
        table["batch_P500_K16"][str(bs)] = avg_ms(lambda: knn_points(x, x, lengths1=L, lengths2=L, K=16, return_nn=False))
    return table


def run_sweep(args):
    from pytorch3d_pointops_b200.functions import ball_query, knn_points, sample_farthest_points

    dev = torch.device("cuda", 0)
    ours = sweep_table(knn_points, ball_query, sample_farthest_points, dev)
    out = {"impl": "ours", "sweep": ours, "unit": "ms per call (wall clock, synchronised per call, 10 runs after 3 warm-ups: "
                                                   "the reference harness's protocol)"}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference_cuda", "--sweep"],
                           capture_output=True, text=True, timeout=900)
        ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        if "sweep" in ref:
            out["reference_cuda"] = ref["sweep"]
            out["speedup"] = {t: {k: ref["sweep"][t][k] / v for k, v in rows.items()} for t, rows in ours.items()}
        else:
            out["reference_cuda"] = ref
    except Exception as e:  # noqa: BLE001
        out["reference_cuda"] = {"unavailable": str(e)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--sweep", action="store_true", help="size sweeps of the reference's timing harness instead of the headline line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-highdim", action="store_true", help="skip the D=128 / FPS / ball-query secondary measurements")
    ap.add_argument("--no-hbm", action="store_true", help="skip the gather / packing / backward secondary measurements")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference CUDA subprocess (vs_reference_cuda)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    elif args.impl == "reference_cuda":
        if rank == 0:
            run_reference_cuda(args)
    elif args.sweep:
        if rank == 0:
            run_sweep(args)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
