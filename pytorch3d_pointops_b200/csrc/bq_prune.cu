// Ball query, D = 3, over Hilbert-ordered clouds: find EVERY point inside the ball through the
// block boxes, then keep the K smallest indices.
//
// Contract = ball_query_cpu.cpp:12-54 of the reference (the FIRST K points of p2, in index order,
// whose unfused float32 squared distance is strictly below radius^2; idx padded with -1, dists with
// 0).  The reference -- and ball_query_scan_kernel here -- walk the cloud in index order and stop at
// K hits, which on sparse balls (expected hits ~ K, the PointNet++ set-abstraction regime) means
// most of the cloud for every query: a CTA leaves the scan only when its slowest query is complete.
// "The first K in index order" is also "the K smallest indices among all hits", and all hits are
// few and close together in space:
//
//   pre-pass (knn_order.cu)   the KNN search's: both clouds along a Hilbert curve, p2 in blocks of 64
//                             points (rows x, y, z, w = |p|^2, original index) with one box each.
//   this kernel               one warp = 32 consecutive sorted queries.
//     walk    every block whose box lies within radius of the warp's query box (exact lower bound of
//             the reference distance: knn_prune_common.cuh); blocks arrive by TMA bulk copies into the
//             warp's ring, as in knn_prune.cu, and are scanned run by run: a run of 16 points only if
//             some query's own ball reaches the run's box.
//     scan    the KNN scan: expanded-form filter against (r^2 - |q|^2) + E, group ids appended,
//             predicated, to the query's candidate buffer.
//     flush   buffered groups re-read from the sorted blocks (L2), exact unfused distance, strict
//             d < r^2: the point's ORIGINAL index goes to the query's hit column in shared memory.
//             A column that fills up (128 hits) is cut back to its K smallest indices by a warp-wide
//             bitonic sort, and from then on only indices below the K-th are admitted.
//     select  per query, the warp sorts the column (32 / 64 / 128-key bitonic network over shuffles);
//     write   lanes = queries again: exact distance of each kept index from the original p2 row
//             (the same unfused operations as the reference, so the same bits), idx / dists rows.
//
// The work is proportional to the points NEAR a query instead of the cloud size; it grows with the
// number of hits, where the index-order scan shrinks (it stops after ~K/hits of the cloud).  Each
// cloud therefore takes one of the two kernels, decided on the device by bq_decide_kernel: it counts the
// hits of 16 sample queries (spread along the curve) and compares two cost models fitted on B200.
#include <cfloat>

#include "bq_prune.cuh"
#include "knn_prune_common.cuh"

namespace pops {

namespace {

constexpr int kBqHitCap = 100;       // hits a query holds between two cuts (>= K + 8).  With 3 ring slots a CTA takes
                                     // 23.2 KB (16-bit ids): 8 CTAs per SM inside the 196 KB carve-out (128 hits and 4
                                     // slots: 31.2 KB, 7 CTAs, 228 KB carve-out = 28 KB of L1)
constexpr int kBqRingSlots = 3;      // blocks resident per warp, kBqRingSlots - 1 in flight ahead of the scan
constexpr int kBqCandCap = 24;       // candidate groups a query buffers between flushes
constexpr int kBqThreadsP = 64;      // two independent warps per CTA

struct BqPruneParams {
  const float4* qsorted;
  const float* blocks;
  const float4* boxes;
  const unsigned* flags;
  const unsigned* maxabs_bits;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  int64_t* idx;
  float* dists;
  int P1, P2, K, nbox;
  float radius, radius2;
  int mode;
};

// IDX: hit index type -- unsigned short while the cloud has at most 65536 points
template <typename CID, typename IDX>
struct BqSmem {
  static constexpr int WARPS = kBqThreadsP / 32;
  static constexpr size_t bars_off = 0;
  static constexpr size_t ring_off = 128;
  static constexpr size_t ring_bytes = size_t(WARPS) * kBqRingSlots * kBlockBytes;
  static constexpr size_t cand_off = ring_off + ring_bytes;
  static constexpr size_t cand_bytes = size_t(kBqCandCap) * kBqThreadsP * sizeof(CID);
  static constexpr size_t hits_off = (cand_off + cand_bytes + 15) / 16 * 16;
  // one column per query, contiguous, padded by one word: lanes appending at similar depths and a warp
  // reading one column both touch 32 different banks
  static constexpr size_t col_bytes = size_t(kBqHitCap) * sizeof(IDX) + 4;
  static constexpr size_t hits_bytes = size_t(kBqThreadsP) * col_bytes;
  static constexpr size_t total = hits_off + hits_bytes;
};

// ascending bitonic sort of 32 * NPL keys held NPL per lane (key e = lane * NPL + r)
template <int NPL>
__device__ __forceinline__ void warp_sort_u32(unsigned (&v)[NPL], int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= NPL) {
        const int lj = j / NPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const unsigned o = __shfl_xor_sync(FULL, v[r], lj);
          const bool up = ((lane * NPL + r) & k) == 0;  // ascending run
          v[r] = (up == lower) ? min(v[r], o) : max(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const int p = r ^ j;
          if (p > r) {
            const bool up = ((lane * NPL + r) & k) == 0;
            const unsigned lo = min(v[r], v[p]), hi = max(v[r], v[p]);
            v[r] = up ? lo : hi;
            v[p] = up ? hi : lo;
          }
        }
      }
    }
  }
}

// Same network on two independent sets at once: set A in the low halves, set B in the high halves of the
// registers (VIMNMX.U16x2: one instruction takes both minima or both maxima).
template <int NPL>
__device__ __forceinline__ void warp_sort_u16x2(unsigned (&v)[NPL], int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= NPL) {
        const int lj = j / NPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const unsigned o = __shfl_xor_sync(FULL, v[r], lj);
          const bool up = ((lane * NPL + r) & k) == 0;
          v[r] = (up == lower) ? __vminu2(v[r], o) : __vmaxu2(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const int p = r ^ j;
          if (p > r) {
            const bool up = ((lane * NPL + r) & k) == 0;
            const unsigned lo = __vminu2(v[r], v[p]), hi = __vmaxu2(v[r], v[p]);
            v[r] = up ? lo : hi;
            v[p] = up ? hi : lo;
          }
        }
      }
    }
  }
}

// Warp-wide, 16-bit indices: sort the columns of TWO queries in one pass of the network.
template <int NPL>
__device__ __forceinline__ void sort_column_pair(unsigned short* ca, int na, unsigned short* cb, int nb, int keep,
                                                 int lane) {
  unsigned v[NPL];
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = lane * NPL + r;
    const unsigned a = e < na ? static_cast<unsigned>(ca[e]) : 0xFFFFu;  // (a real index 65535 ties with the padding:
    const unsigned b = e < nb ? static_cast<unsigned>(cb[e]) : 0xFFFFu;  //  equal values, either order is the same)
    v[r] = a | (b << 16);
  }
  __syncwarp();
  warp_sort_u16x2<NPL>(v, lane);
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = lane * NPL + r;
    if (e < na && e < keep) ca[e] = static_cast<unsigned short>(v[r] & 0xFFFFu);
    if (e < nb && e < keep) cb[e] = static_cast<unsigned short>(v[r] >> 16);
  }
  __syncwarp();
}

__device__ __noinline__ void sort_column_pair_any(unsigned short* ca, int na, unsigned short* cb, int nb, int keep,
                                                  int lane) {
  const int n = max(na, nb);
  if (n <= 1) return;
  if (n <= 32) sort_column_pair<1>(ca, na, cb, nb, keep, lane);
  else if (n <= 64) sort_column_pair<2>(ca, na, cb, nb, keep, lane);
  else sort_column_pair<4>(ca, na, cb, nb, keep, lane);
}

// Warp-wide: sort the n (<= 32 * NPL) hits of one column, write the smallest min(n, keep) back in
// ascending order.
template <int NPL, typename IDX>
__device__ __forceinline__ void sort_column(IDX* col, int n, int keep, int lane) {
  unsigned v[NPL];
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = lane * NPL + r;
    v[r] = e < n ? static_cast<unsigned>(col[e]) : 0xFFFFFFFFu;
  }
  __syncwarp();
  warp_sort_u32<NPL>(v, lane);
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = lane * NPL + r;
    if (e < n && e < keep) col[e] = static_cast<IDX>(v[r]);
  }
  __syncwarp();
}

template <typename IDX>
__device__ __noinline__ void sort_column_any(IDX* col, int n, int keep, int lane) {
  if (n <= 1) return;
  if (n <= 32) sort_column<1, IDX>(col, n, keep, lane);
  else if (n <= 64) sort_column<2, IDX>(col, n, keep, lane);
  else sort_column<4, IDX>(col, n, keep, lane);
}

// Drain one warp's candidate buffers: lanes = queries.  Groups are re-read from the sorted blocks (L2),
// two per round for memory-level parallelism; a point is a hit if its exact distance is below r2 and
// its index below the lane's admission bound.  Between rounds a column that could overflow in the next
// one is cut back to its K smallest indices (warp-wide sort), which also lowers the bound.
template <typename CID, typename IDX>
__device__ __noinline__ void bq_flush(const float* __restrict__ blocks_n, const CID* cand_col, int c_end,
                                      IDX* hits_warp, size_t col_stride_idx, float q0, float q1, float q2, float r2,
                                      int K, int& n_io, unsigned& imax_io) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  IDX* col = hits_warp + static_cast<size_t>(lane) * col_stride_idx;
  int n = n_io;
  unsigned imax = imax_io;
  int c = 0;
  while (__any_sync(FULL, c < c_end)) {
    if (c < c_end) {
      const bool two = c + 1 < c_end;
      const unsigned ga = cand_col[c * kBqThreadsP];
      const unsigned gb = two ? cand_col[(c + 1) * kBqThreadsP] : ga;
      c += 2;
      const float* pa = blocks_n + static_cast<size_t>(ga / kBlockGroups) * kBlockFloats + (ga % kBlockGroups) * kGroup;
      const float* pb = blocks_n + static_cast<size_t>(gb / kBlockGroups) * kBlockFloats + (gb % kBlockGroups) * kGroup;
      const float4 Xa = *reinterpret_cast<const float4*>(pa);
      const float4 Ya = *reinterpret_cast<const float4*>(pa + kBoxPoints);
      const float4 Za = *reinterpret_cast<const float4*>(pa + 2 * kBoxPoints);
      const uint4 Ia = *reinterpret_cast<const uint4*>(pa + kIdxOff);
      const float4 Xb = *reinterpret_cast<const float4*>(pb);
      const float4 Yb = *reinterpret_cast<const float4*>(pb + kBoxPoints);
      const float4 Zb = *reinterpret_cast<const float4*>(pb + 2 * kBoxPoints);
      const uint4 Ib = *reinterpret_cast<const uint4*>(pb + kIdxOff);
      float da[4], db[4];
      exact4(q0, q1, q2, Xa, Ya, Za, da);
      exact4(q0, q1, q2, Xb, Yb, Zb, db);
      const unsigned ia[4] = {Ia.x, Ia.y, Ia.z, Ia.w}, ib[4] = {Ib.x, Ib.y, Ib.z, Ib.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (da[i] < r2 && ia[i] < imax) {  // padding entries carry index 0xFFFFFFFF >= imax
          col[n] = static_cast<IDX>(ia[i]);
          ++n;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (two && db[i] < r2 && ib[i] < imax) {
          col[n] = static_cast<IDX>(ib[i]);
          ++n;
        }
      }
    }
    // columns that could not take another round (8 hits): keep their K smallest indices
    unsigned full = __ballot_sync(FULL, n > kBqHitCap - 2 * kGroup);
    while (full) {
      const int ql = __ffs(full) - 1;
      full &= full - 1;
      const int nq = __shfl_sync(FULL, n, ql);
      IDX* cq = hits_warp + static_cast<size_t>(ql) * col_stride_idx;
      __syncwarp();
      sort_column_any<IDX>(cq, nq, K, lane);
      if (lane == ql && nq >= K) {
        n = K;
        imax = static_cast<unsigned>(cq[K - 1]);  // a later hit counts only if it displaces the K-th
      }
    }
  }
  n_io = n;
  imax_io = imax;
}

template <typename CID, typename IDX>
__global__ void __launch_bounds__(kBqThreadsP, 8)
bq_prune_kernel(const BqPruneParams prm) {
  constexpr int S = kBqRingSlots, THREADS = kBqThreadsP;
  using SM = BqSmem<CID, IDX>;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr unsigned FULL = 0xffffffffu;

  const int n = blockIdx.y;
  const int q_base = blockIdx.x * THREADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = prm.K;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  if (prm.flags[n] == 0u) return;  // the index-order scan answers this cloud
  int64_t* out_idx = prm.idx + (static_cast<size_t>(n) * prm.P1) * K;
  float* out_d = prm.dists + (static_cast<size_t>(n) * prm.P1) * K;
  const float4* qs = prm.qsorted + static_cast<size_t>(n) * prm.P1;
  const float INF = __int_as_float(0x7f800000);
  const float r2 = prm.radius2;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::bars_off) + warp * S;
  float* ring = reinterpret_cast<float*>(smem + SM::ring_off) + static_cast<size_t>(warp) * S * kScanFloats;
  const float4* ring4 = reinterpret_cast<const float4*>(ring);
  CID* cand = reinterpret_cast<CID*>(smem + SM::cand_off);
  constexpr size_t COLI = SM::col_bytes / sizeof(IDX);  // column stride in index elements
  IDX* hits_warp = reinterpret_cast<IDX*>(smem + SM::hits_off) + static_cast<size_t>(warp) * 32 * COLI;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  // ---- the lane's query ---------------------------------------------------------------------------
  const int qi = q_base + tid;
  const int wq0 = q_base + warp * 32;
  const bool inrow = qi < prm.P1;
  const bool valid = qi < L1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (inrow) v = qs[qi];
  const unsigned row = __float_as_uint(v.w);
  const float q0 = valid ? v.x : 0.0f, q1 = valid ? v.y : 0.0f, q2 = valid ? v.z : 0.0f;
  const float a0 = -2.0f * q0, a1 = -2.0f * q1, a2 = -2.0f * q2;
  const float qq = fmaf(q2, q2, fmaf(q1, q1, q0 * q0));
  const float M = __uint_as_float(prm.maxabs_bits[n]);
  const float E = fmaf(M * M, 1.52587890625e-05f /* 2^-16 */, 1e-37f);  // filter error bound
  const float T = valid ? __fadd_rn(__fsub_rn(r2, qq), E) : -INF;
  int nhits = 0;
  unsigned imax = 0xFFFFFFFFu;  // indices at or above are not admitted (padding entries: 0xFFFFFFFF)

  const int nblk = (L2 + kBoxPoints - 1) / kBoxPoints;
  if (wq0 < L1 && nblk > 0) {
    float wqlo[3] = {valid ? q0 : INF, valid ? q1 : INF, valid ? q2 : INF};
    float wqhi[3] = {valid ? q0 : -INF, valid ? q1 : -INF, valid ? q2 : -INF};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        wqlo[d] = fminf(wqlo[d], __shfl_xor_sync(FULL, wqlo[d], o));
        wqhi[d] = fmaxf(wqhi[d], __shfl_xor_sync(FULL, wqhi[d], o));
      }
    }
    const float qown[3] = {q0, q1, q2};

    // ---- block walk: every block within the radius of the warp's query box, in block order ---------
    const int nchunks = (nblk + 31) >> 5;
    const float4* boxes_n = prm.boxes + static_cast<size_t>(n) * prm.nbox * 2;
    int ch = -1;
    unsigned pend = 0u;
    auto pick = [&]() -> int {
      for (;;) {
        if (pend) {
          const int b = __ffs(pend) - 1;
          pend &= pend - 1;
          return (ch << 5) + b;
        }
        if (++ch >= nchunks) return -1;
        const float4* bx = boxes_n + static_cast<size_t>(ch * 32 + lane) * 2;  // nbox is a multiple of 32
        const float lb = box_lower_bound(bx[0], bx[1], wqlo, wqhi);
        const int left = nblk - (ch << 5);
        pend = __ballot_sync(FULL, lb < r2) & (left >= 32 ? FULL : ((1u << left) - 1u));
      }
    };

    const float* blocks_n = prm.blocks + static_cast<size_t>(n) * prm.nbox * kBlockFloats;
    int slot_blk = 0;  // lane s: block in ring slot s
    int head = 0, tail = 0;
    auto issue = [&](int b) {
      const int s = head % S;
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();  // the slot's previous contents were read through the generic proxy
        mbar_arrive_expect_tx(&bars[s], kBlockBytes);
        tma_bulk_g2s(ring + static_cast<size_t>(s) * kScanFloats, blocks_n + static_cast<size_t>(b) * kBlockFloats,
                     kBlockBytes, &bars[s]);
      }
      if (lane == s) slot_blk = b;
      ++head;
    };

    constexpr uint32_t CB = sizeof(CID);
    constexpr uint32_t CBYTES = THREADS * CB;
    const uint32_t cand_base = smem_u32(cand) + static_cast<uint32_t>(tid) * CB;
    const uint32_t cw_limit = cand_base + static_cast<uint32_t>(kBqCandCap - kChunk) * CBYTES;
    uint32_t cw = cand_base;
    auto flush = [&]() {
      const int c_end = static_cast<int>((cw - cand_base) / CBYTES);
      cw = cand_base;
      bq_flush<CID, IDX>(blocks_n, cand + tid, c_end, hits_warp, COLI, q0, q1, q2, r2, K, nhits, imax);
    };

    for (;;) {
      while (head - tail < S - 1) {
        const int b = pick();
        if (b < 0) break;
        issue(b);
      }
      if (tail == head) break;
      const int s = tail % S;
      mbar_wait(&bars[s], (tail / S) & 1);
      // the warp-wide test used the box of ALL its queries: a run of kSubPoints points of the block is scanned
      // only if some query's own ball reaches the run's box (the boxes arrive with the block, knn_order.cuh)
      static_assert(kChunk * kGroup == kSubPoints && kSubBoxes == 4, "one overflow-check chunk of the scan = one run");
      const float4* sb = ring4 + s * kBlockF4 + kSubOff / 4;
      unsigned sub = 0u;
#pragma unroll
      for (int j = 0; j < kSubBoxes; ++j) {
        const bool need = valid && box_lower_bound(sb[2 * j], sb[2 * j + 1], qown, qown) < r2;
        sub |= __any_sync(FULL, need) ? (1u << j) : 0u;
      }
      if (sub) {
        const float4* tp = ring4 + s * kBlockF4;
        const unsigned gid0 = static_cast<unsigned>(__shfl_sync(FULL, slot_blk, s)) * kBlockGroups;
        do {  // (the flush is a call and stays outside the dense loop, as in knn_prune.cu)
          int g = (__ffs(sub) - 1) * kChunk;  // first run still to scan
          unsigned gid = gid0 + static_cast<unsigned>(g);
          float4 Xc[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) Xc[r] = tp[r * kBlockGroups + g];
          bool over = false;
#pragma unroll 1
          for (; g < kBlockGroups && !over && ((sub >> (g / kChunk)) & 1u); g += kChunk) {
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
              float4 Xn[4];  // next group's rows (the last prefetch of a block reads the run boxes: in bounds, unused)
#pragma unroll
              for (int r = 0; r < 4; ++r) Xn[r] = tp[r * kBlockGroups + g + c + 1];
              float2 s01 = make_float2(Xc[3].x, Xc[3].y), s23 = make_float2(Xc[3].z, Xc[3].w);
              s01 = __ffma2_rn(make_float2(a0, a0), make_float2(Xc[0].x, Xc[0].y), s01);
              s23 = __ffma2_rn(make_float2(a0, a0), make_float2(Xc[0].z, Xc[0].w), s23);
              s01 = __ffma2_rn(make_float2(a1, a1), make_float2(Xc[1].x, Xc[1].y), s01);
              s23 = __ffma2_rn(make_float2(a1, a1), make_float2(Xc[1].z, Xc[1].w), s23);
              s01 = __ffma2_rn(make_float2(a2, a2), make_float2(Xc[2].x, Xc[2].y), s01);
              s23 = __ffma2_rn(make_float2(a2, a2), make_float2(Xc[2].z, Xc[2].w), s23);
              const float m = fminf(fminf(s01.x, s01.y), fminf(s23.x, s23.y));
              if (m <= T) {  // predicated: one STS + one IADD
                if (sizeof(CID) == 2)
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(cw), "h"(static_cast<unsigned short>(gid)) : "memory");
                else
                  asm volatile("st.shared.u32 [%0], %1;" ::"r"(cw), "r"(gid) : "memory");
                cw += CBYTES;
              }
              ++gid;
#pragma unroll
              for (int r = 0; r < 4; ++r) Xc[r] = Xn[r];
            }
            over = __any_sync(FULL, cw > cw_limit);
          }
          sub &= ~((1u << (g / kChunk)) - 1u);  // runs below g are done or were skipped
          if (over) flush();
        } while (sub);
      }
      ++tail;
    }
    flush();
  }

  // ---- select: the warp sorts every query's column; the first min(hits, K) indices are the answer ----
  if constexpr (sizeof(IDX) == 2) {  // two queries per pass of the network (packed 16-bit minima / maxima)
    for (int ql = 0; ql < 32; ql += 2) {
      const int na = __shfl_sync(FULL, nhits, ql), nb = __shfl_sync(FULL, nhits, ql + 1);
      sort_column_pair_any(reinterpret_cast<unsigned short*>(hits_warp) + static_cast<size_t>(ql) * COLI, na,
                           reinterpret_cast<unsigned short*>(hits_warp) + static_cast<size_t>(ql + 1) * COLI, nb, K, lane);
    }
  } else {
    for (int ql = 0; ql < 32; ++ql) {
      const int nq = __shfl_sync(FULL, nhits, ql);
      if (nq > 1) sort_column_any<IDX>(hits_warp + static_cast<size_t>(ql) * COLI, nq, K, lane);
    }
  }
  __syncwarp();

  // ---- write: lanes = queries; distances from the original rows, the reference's own operations ------
  if (!inrow) return;
  const IDX* col = hits_warp + static_cast<size_t>(lane) * COLI;
  const int kept = min(nhits, K);
  int64_t* oi = out_idx + static_cast<size_t>(row) * K;
  float* od = out_d + static_cast<size_t>(row) * K;
  const float* p2n = prm.p2 + static_cast<size_t>(n) * prm.P2 * 3;
  int k = 0;
  for (; k + 4 <= kept; k += 4) {
    unsigned j[4];
    float x[4], y[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      j[u] = static_cast<unsigned>(col[k + u]);
      const float* p = p2n + static_cast<size_t>(j[u]) * 3;
      x[u] = p[0]; y[u] = p[1]; z[u] = p[2];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float dx = __fsub_rn(q0, x[u]), dy = __fsub_rn(q1, y[u]), dz = __fsub_rn(q2, z[u]);
      od[k + u] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      oi[k + u] = static_cast<int64_t>(j[u]);
    }
  }
  for (; k < kept; ++k) {
    const unsigned j = static_cast<unsigned>(col[k]);
    const float* p = p2n + static_cast<size_t>(j) * 3;
    const float dx = __fsub_rn(q0, p[0]), dy = __fsub_rn(q1, p[1]), dz = __fsub_rn(q2, p[2]);
    od[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    oi[k] = static_cast<int64_t>(j);
  }
  for (; k < K; ++k) {
    oi[k] = -1;
    od[k] = 0.0f;
  }
}

// One CTA per cloud: hits of kBqSamples sample queries (sorted queries at regular positions along the
// curve) over the whole cloud, by the expanded form -- an estimate is all that is needed.  Then
//   cost of finding all hits  ~ 4900 + 49 * mean hits            (point evaluations per query)
//   cost of the index scan    ~ L2 * min(1, 0.2 + 1.6 K / fewest hits)
// (a CTA of the scan runs until its slowest query holds K hits, so the sparsest ball sets its time; the
// constants were fitted on 32 x 16384 uniform and spherical-shell clouds, K = 16 / 32 / 64, r = 0.05 .. 0.3).
constexpr int kBqSamples = 16;
constexpr int kBqDecideThreads = 512;

__global__ void __launch_bounds__(kBqDecideThreads)
bq_decide_kernel(const float4* __restrict__ qsorted, const float* __restrict__ p2, const int64_t* __restrict__ len1,
                 const int64_t* __restrict__ len2, const unsigned* __restrict__ maxabs_bits, int P1, int P2, int K,
                 float radius, float radius2, int mode, unsigned* __restrict__ flags) {
  __shared__ float sq[kBqSamples][4];  // -2 q, r2 - |q|^2
  __shared__ int total[kBqSamples];
  const int n = blockIdx.x, tid = threadIdx.x;
  int64_t L1l = len1[n], L2l = len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > P2 ? P2 : L2l));
  const float r = fabsf(radius);
  bool fixed = false, take = false;  // decided without counting
  if (maxabs_bits[n] >= kDirtyBits || !(r > 0.0f) || !(r < 1e18f) || L1 == 0 || L2 == 0) fixed = true;  // exact scan
  else if (mode == 1) fixed = take = true;
  else if (L2 < kBqSpatialMinPoints) fixed = true;
  if (fixed) {
    if (tid == 0) flags[n] = take ? 1u : 0u;
    return;
  }
  if (tid < kBqSamples) {
    const int pos = static_cast<int>((static_cast<int64_t>(2 * tid + 1) * L1) / (2 * kBqSamples));
    const float4 q = qsorted[static_cast<size_t>(n) * P1 + pos];
    sq[tid][0] = -2.0f * q.x; sq[tid][1] = -2.0f * q.y; sq[tid][2] = -2.0f * q.z;
    sq[tid][3] = radius2 - fmaf(q.z, q.z, fmaf(q.y, q.y, q.x * q.x));
    total[tid] = 0;
  }
  __syncthreads();
  int cnt[kBqSamples];
#pragma unroll
  for (int s = 0; s < kBqSamples; ++s) cnt[s] = 0;
  const float* pts = p2 + static_cast<size_t>(n) * P2 * 3;
  for (int j = tid; j < L2; j += kBqDecideThreads) {
    const float x = pts[static_cast<size_t>(j) * 3], y = pts[static_cast<size_t>(j) * 3 + 1], z = pts[static_cast<size_t>(j) * 3 + 2];
    const float w = fmaf(z, z, fmaf(y, y, x * x));
#pragma unroll
    for (int s = 0; s < kBqSamples; ++s) {
      const float e = fmaf(sq[s][2], z, fmaf(sq[s][1], y, fmaf(sq[s][0], x, w)));
      cnt[s] += e < sq[s][3] ? 1 : 0;
    }
  }
#pragma unroll
  for (int s = 0; s < kBqSamples; ++s) {
    const int c = __reduce_add_sync(0xffffffffu, cnt[s]);
    if ((tid & 31) == 0 && c) atomicAdd(&total[s], c);
  }
  __syncthreads();
  if (tid == 0) {
    int sum = 0, mn = 0x7fffffff;
    for (int s = 0; s < kBqSamples; ++s) {
      sum += total[s];
      mn = min(mn, total[s]);
    }
    const float mean = static_cast<float>(sum) / kBqSamples;
    const float all_hits = 4900.0f + 49.0f * mean;
    const float scan = static_cast<float>(L2) * fminf(1.0f, 0.2f + 1.6f * static_cast<float>(K) / fmaxf(static_cast<float>(mn), 1.0f));
    flags[n] = all_hits < scan ? 1u : 0u;
  }
}

template <typename CID, typename IDX>
int launch_bq_prune(const BqPruneParams& prm, int N, cudaStream_t st) {
  using SM = BqSmem<CID, IDX>;
  auto kern = bq_prune_kernel<CID, IDX>;
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::total)));
  dim3 grid(static_cast<unsigned>(ceil_div(prm.P1, kBqThreadsP)), N);
  kern<<<grid, kBqThreadsP, SM::total, st>>>(prm);
  POPS_LAUNCH_OK("bq_prune_kernel");
  return POPS_OK;
}

}  // namespace

int bq_prune_search(const KnnOrderBuffers& ob, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                    int P1, int P2, int K, float radius, float radius2, int mode, unsigned* flags, int64_t* idx,
                    float* dists, cudaStream_t st) {
  bq_decide_kernel<<<static_cast<unsigned>(N), kBqDecideThreads, 0, st>>>(ob.qsorted, p2, len1, len2, ob.maxabs_bits, P1,
                                                                         P2, K, radius, radius2, mode, flags);
  POPS_LAUNCH_OK("bq_decide_kernel");
  BqPruneParams prm;
  prm.qsorted = ob.qsorted; prm.blocks = ob.blocks; prm.boxes = ob.boxes; prm.flags = flags;
  prm.maxabs_bits = ob.maxabs_bits; prm.p2 = p2; prm.len1 = len1; prm.len2 = len2; prm.idx = idx; prm.dists = dists;
  prm.P1 = P1; prm.P2 = P2; prm.K = K; prm.nbox = static_cast<int>(knn_order_num_boxes(P2));
  prm.radius = radius; prm.radius2 = radius2; prm.mode = mode;
  if (P2 <= 65536) return launch_bq_prune<unsigned short, unsigned short>(prm, N, st);
  const bool narrow = int64_t(prm.nbox) * kBlockGroups <= 65536;  // group ids fit 16 bits
  return narrow ? launch_bq_prune<unsigned short, unsigned>(prm, N, st) : launch_bq_prune<unsigned, unsigned>(prm, N, st);
}

}  // namespace pops
