"""Host-resident KNN: clouds stream through the GPU in slices.

`knn_points` (functions/knn.py, the reference's API) takes device tensors.  When the clouds live
in host memory and the (idx, dists) result is wanted back on the host -- 100 MB for the
B=32 x P=16384 x K=16 shape, most of the end-to-end time -- the batch is independent per cloud
(outer `for n` of knn_cpu.cpp:35), so the legs pipeline: the first slice goes in, is ordered and
searched on its own; the rest of the batch follows with one copy and ONE spatial pre-pass, and
while slice i is searched the results of slice i-1 are on their way out.  Copies run on two side streams, the kernels on the caller's current
stream; buffers are pinned once and reused.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _C


class HostKnn:
    """Reusable pinned staging + streams for host-in / host-out knn_points on one device.

    out_idx (N,P1,K) int64 and out_dists (N,P1,K) float32 are pinned host tensors owned by this
    object and overwritten by every call."""

    def __init__(self, N: int, P1: int, P2: int, D: int, K: int, device, slices=8, graph: bool = True,
                 idx_dtype: torch.dtype = torch.int64):
        # idx_dtype: dtype of the HOST index buffer.  int64 is the reference's contract (knn.h:59-66);
        # int32 (additive) narrows each slice on the device before it travels -- a third less D2H
        # traffic, which is what bounds this pipeline -- for callers that widen on the host or do not
        # need 64-bit indices (P2 < 2^31 always holds).
        assert idx_dtype in (torch.int64, torch.int32)
        self.idx_dtype = idx_dtype
        self.device = torch.device(device)
        self.N, self.P1, self.P2, self.D, self.K = N, P1, P2, D, K
        # `slices`: a count (equal slices) or an explicit list of slice sizes in clouds.  Measured on
        # the B=32 x P=16384 x K=16 shape (D2H of the 100 MB result alone: 1.78 ms): 2 slices 2.47 ms,
        # 4: 2.18, 6: 2.16, 8: 2.11, 10: 2.15 -- a slice below one wave of CTAs still costs one CTA's
        # latency, so many slices stretch the search; uneven slice sizes and searching the slices on 2-3
        # alternating streams end at the same floor.
        if isinstance(slices, (list, tuple)):
            sizes = [int(v) for v in slices if int(v) > 0]
            assert sum(sizes) == N, "slice sizes must add up to the batch"
            bounds = [0]
            for v in sizes:
                bounds.append(bounds[-1] + v)
            self.slices = len(sizes)
        else:
            self.slices = max(1, min(int(slices), N))
            bounds = [round(i * N / self.slices) for i in range(self.slices + 1)]
        self.out_idx = torch.empty((N, P1, K), dtype=idx_dtype).pin_memory()
        self.out_dists = torch.empty((N, P1, K), dtype=torch.float32).pin_memory()
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self.ranges = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        # The whole pipeline (copies on two side streams, the pre-pass, one search per slice) is captured into ONE
        # CUDA graph per set of host buffers and replayed: the slices are short enough that
        # launching them from Python would leave the GPU idle between them.
        self.use_graph = graph
        self._graph_key = None
        self._graph = None
        self._keep = None

    def __call__(self, p1: torch.Tensor, p2: Optional[torch.Tensor] = None,
                 lengths1: Optional[torch.Tensor] = None, lengths2: Optional[torch.Tensor] = None,
                 norm: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
        """p1 (N,P1,D), p2 (N,P2,D) [None: self-KNN], lengths (N,) int64: pinned HOST tensors.
        Returns (out_dists, out_idx) -- valid once the current stream has been synchronised."""
        self_knn = p2 is None
        if lengths1 is None:
            lengths1 = self._full_lengths(self.P1)
        if lengths2 is None:
            lengths2 = lengths1 if self_knn else self._full_lengths(self.P2)
        if not self.use_graph:
            self._run(p1, p2, lengths1, lengths2, norm, recording=False)
            return self.out_dists, self.out_idx
        key = (p1.data_ptr(), 0 if self_knn else p2.data_ptr(), lengths1.data_ptr(), lengths2.data_ptr(), norm)
        if key != self._graph_key:
            torch.cuda.synchronize(self.device)
            self._run(p1, p2, lengths1, lengths2, norm, recording=False)  # warm-up outside the capture
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._keep = self._run(p1, p2, lengths1, lengths2, norm, recording=True)
            self._graph, self._graph_key = g, key
            self._inputs = (p1, p2, lengths1, lengths2)  # the graph reads these host buffers: keep them alive
        self._graph.replay()
        return self.out_dists, self.out_idx

    def _full_lengths(self, P):
        cache = self.__dict__.setdefault("_full", {})
        if P not in cache:
            cache[P] = torch.full((self.N,), P, dtype=torch.int64).pin_memory()
        return cache[P]

    def _run(self, p1, p2, lengths1, lengths2, norm, recording):
        dev = self.device
        main = torch.cuda.current_stream(dev)
        self_knn = p2 is None
        same_len = self_knn and lengths2 is lengths1
        # Two groups of slices: the FIRST slice travels, is ordered and searched on its own, so that its
        # results are on their way back while the rest of the batch is still going in; the remaining
        # slices share one copy and one pre-pass (pointops_b200.h: pops_knn_points_prepare /
        # pops_knn_points_idx_range).
        groups = [self.ranges[:1], self.ranges[1:]] if len(self.ranges) > 2 else [self.ranges]
        staged = []
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_stream(main)
            for grp in groups:
                g0, g1 = grp[0][0], grp[-1][1]
                d1 = p1[g0:g1].to(dev, non_blocking=True)
                d2 = d1 if self_knn else p2[g0:g1].to(dev, non_blocking=True)
                l1 = lengths1[g0:g1].to(dev, non_blocking=True)
                l2 = l1 if same_len else lengths2[g0:g1].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.h2d)
                staged.append((d1, d2, l1, l2, ev))
        keep, narrow = [], []
        for grp, (d1, d2, l1, l2, ev) in zip(groups, staged):
            g0 = grp[0][0]
            main.wait_event(ev)
            if not recording:
                for t in (d1, d2, l1, l2):
                    t.record_stream(main)
            ks = _C.KnnSliced(d1, d2, l1, l2, norm, self.K)
            ks.prepare()
            for a, b in grp:
                ks.search(a - g0, b - g0)
                idx_src = ks.idx[a - g0:b - g0]
                if self.idx_dtype != torch.int64:
                    idx_src = idx_src.to(self.idx_dtype)  # narrowed on the device, on the search stream
                    if not recording:
                        idx_src.record_stream(self.d2h)
                    narrow.append(idx_src)
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    self.out_idx[a:b].copy_(idx_src, non_blocking=True)
                    self.out_dists[a:b].copy_(ks.dists[a - g0:b - g0], non_blocking=True)
            if not recording:
                for t in (ks.idx, ks.dists, ks.ws):
                    t.record_stream(self.d2h)
            keep.append((d1, d2, l1, l2, ks))
        keep.append(narrow)
        main.wait_stream(self.d2h)
        return keep
