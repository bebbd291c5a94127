// Pieces of the Hilbert-ordered, box-pruned D = 3 search (knn_prune.cu) that do not depend on its
// template parameters: launch parameters, the exact lower bound between boxes, the exact
// distance of one query to a group of four points, and the seed bound.
#pragma once
#include <cfloat>

#include "knn_core.cuh"

namespace pops {

struct KnnPruneParams {
  const float4* qsorted;
  const unsigned* qhome;
  const float* blocks;
  const float4* boxes;
  const int64_t* len1;
  const int64_t* len2;
  const unsigned* maxabs_bits;
  int64_t* idx;
  float* dists;
  int P1, P2, K, nbox;
  int prune;  // 0: visit every block (brute force in the same order); measurement aid
  int subq;   // 1: per-query sub-box test for every K
  int nseed, bufcap;  // tuning aids (0: defaults)
  unsigned long long* stats;  // development counters (POPS_KNN_STATS=1), else nullptr
};

namespace {

constexpr int kRingSlots = 4;   // blocks resident per warp (ball query; KNN: prune_ring_slots)
constexpr int kPrefetch = 3;    // blocks in flight ahead of the scan
// KNN: K <= 16 keeps 3 blocks per warp (2 in flight ahead of the scan).  With the index row out of the ring slot
// that is 19.2 KB of shared memory per CTA: 8 CTAs fit the 164 KB carve-out and the L1 grows from 60 to 92 KB
// (the flush's re-reads of buffered groups and the list rows live there).  K = 32 seeds from 4 blocks.
constexpr int prune_ring_slots(int KT) { return KT <= 16 ? 3 : 4; }
// candidate groups a query can buffer between flushes (a query meets ~K/4 + curve scatter groups in total).
// Larger buffers mean fewer flush rounds (K = 16: 24 -> 28 -> 32 entries: 614 -> 593 -> ... us), as long as the CTA
// stays inside its shared-memory carve-out: 48 entries (3 KB more) pushed 8 CTAs from the 196 KB into the 228 KB
// carve-out, L1 dropped from 60 to 28 KB and the kernel lost 9 % (the flush's re-reads and the list rows live
// in L1).  K = 16 measured 576 us with 32 entries (19.8 KB per CTA, 8 CTAs in the 164 KB carve-out, L1 = 92 KB),
// 560 us with 52 or 64 (22.4 / 23.9 KB, 8 CTAs in the 196 KB carve-out, L1 = 60 KB: ncu reports 8 resident
// CTAs) and 573 us with 84 (228 KB carve-out again): it takes 52.  An earlier FIRST flush only loses.
// (Q = queries per thread: the sizes above are for the forms that run by default, Q = 1 for K <= 16 and Q = 2 for
//  K = 32; the other forms keep round 1's 24 / 40 entries)
constexpr int prune_buf_cap(int KT, int Q) {
  return KT > 16 ? (Q == 2 ? 48 : 40) : (KT == 1 ? 12 : (KT == 16 && Q == 1 ? 52 : 24));
}
constexpr int kBlockF4 = kScanFloats / 4;         // float4 per ring slot: the part of a block a scan reads
constexpr int kBlockGroups = kBoxPoints / kGroup;  // 16 groups of 4 points
constexpr uint32_t kBlockBytes = kScanFloats * 4;  // bytes of one TMA bulk copy (rows x, y, z, w + run boxes)
static_assert((kRingSlots & (kRingSlots - 1)) == 0 && kRingSlots <= 32, "slot metadata sits in lanes");



// CID: candidate id type -- unsigned short while the cloud has at most 65536 groups (262144 points)
template <int Q, int THREADS, typename CID, int KT>
struct PruneSmem {
  static constexpr int WARPS = THREADS / 32;
  static constexpr int QPB = Q * THREADS;
  // Every byte counts: 8 CTAs x (total + 1 KB reserved) must stay within the 164 KB shared-memory carve-out
  // (knn_prune_common.cuh: prune_ring_slots), or at least within the 196 KB one (K = 16 with its 52-entry buffer).
  static constexpr size_t bars_off = 0;      // WARPS x SLOTS mbarriers (<= 64 bytes)
  static constexpr size_t wbox_off = 64;     // per warp: min xyz, -, max xyz, - of its queries (WARPS x 32 bytes)
  static constexpr size_t ring_off = 128;
  static constexpr int SLOTS = prune_ring_slots(KT);
  static constexpr size_t ring_bytes = size_t(WARPS) * SLOTS * kBlockBytes;
  static constexpr size_t cand_off = ring_off + ring_bytes;
  static constexpr size_t cand_bytes = size_t(prune_buf_cap(KT, Q)) * QPB * sizeof(CID);  // global group ids
  static constexpr size_t surv_off = (cand_off + cand_bytes + 15) / 16 * 16;
  static constexpr size_t surv_bytes = size_t(kSurvCap) * THREADS * 8;
  static constexpr size_t cold_off = surv_off + surv_bytes;
  static constexpr size_t cold_bytes = size_t(2) * QPB * 4;  // dk, output row per query
  static constexpr size_t total = cold_off + cold_bytes;
  static_assert(WARPS * SLOTS * 8 <= wbox_off && wbox_off + WARPS * 32 <= ring_off, "mbarriers / query boxes overlap the ring");
};

// Lower bound of the reference distance between ANY query in the box [qlo, qhi] and ANY point in
// the box [lo, hi].  Same unfused operations in the same order as the reference distance
// (knn_cpu.cpp:42-50); rounding is monotone, so for every such pair and every axis
// |fl(q - p)| >= gap, fl(gap^2) <= fl(diff^2), and the rounded sums keep the order: the bound
// holds exactly.  Empty boxes (+inf, -inf) give +inf.
__device__ __forceinline__ float box_lower_bound(float4 lo, float4 hi, const float (&qlo)[3],
                                                 const float (&qhi)[3]) {
  const float gx = fmaxf(fmaxf(__fsub_rn(lo.x, qhi[0]), __fsub_rn(qlo[0], hi.x)), 0.0f);
  const float gy = fmaxf(fmaxf(__fsub_rn(lo.y, qhi[1]), __fsub_rn(qlo[1], hi.y)), 0.0f);
  const float gz = fmaxf(fmaxf(__fsub_rn(lo.z, qhi[2]), __fsub_rn(qlo[2], hi.z)), 0.0f);
  return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
}

// exact unfused distances of q to the 4 points of one group (packed sub / mul, scalar adds:
// ptxas fuses packed mul + packed add into FFMA2, which would break bit parity -- knn_core.cuh)
__device__ __forceinline__ void exact4(float q0, float q1, float q2, float4 X, float4 Y, float4 Z,
                                       float (&d4)[4]) {
  const float2 x01 = __fadd2_rn(make_float2(q0, q0), make_float2(-X.x, -X.y));
  const float2 x23 = __fadd2_rn(make_float2(q0, q0), make_float2(-X.z, -X.w));
  const float2 y01 = __fadd2_rn(make_float2(q1, q1), make_float2(-Y.x, -Y.y));
  const float2 y23 = __fadd2_rn(make_float2(q1, q1), make_float2(-Y.z, -Y.w));
  const float2 z01 = __fadd2_rn(make_float2(q2, q2), make_float2(-Z.x, -Z.y));
  const float2 z23 = __fadd2_rn(make_float2(q2, q2), make_float2(-Z.z, -Z.w));
  const float2 xx01 = __fmul2_rn(x01, x01), xx23 = __fmul2_rn(x23, x23);
  const float2 yy01 = __fmul2_rn(y01, y01), yy23 = __fmul2_rn(y23, y23);
  const float2 zz01 = __fmul2_rn(z01, z01), zz23 = __fmul2_rn(z23, z23);
  d4[0] = __fadd_rn(__fadd_rn(xx01.x, yy01.x), zz01.x);
  d4[1] = __fadd_rn(__fadd_rn(xx01.y, yy01.y), zz01.y);
  d4[2] = __fadd_rn(__fadd_rn(xx23.x, yy23.x), zz23.x);
  d4[3] = __fadd_rn(__fadd_rn(xx23.y, yy23.y), zz23.y);
}

// Seed bound of one query from the `nseed` blocks at the start of the warp's ring.  Every seed point
// gets the EXPANDED form s = w - 2 q.p of the scan (3 FMA per point, a_d = -2 q_d); the minimum of s
// over each of NS interleaved subsets of the points, then the KT-th smallest of those minima, U_s: KT
// distinct points (the subsets are disjoint) have s <= U_s.  By the filter's error bound read the
// other way (|s + qq - d_ref| plus the roundings of forming the sum stay below E), each of
// them has a reference distance d_ref <= fl(fl(U_s + qq) + E), which therefore bounds the K-th
// distance (K <= KT) from above.  +inf when fewer than KT subsets hold a valid point (padding entries
// carry w = +inf).  K = 1 takes the exact distances of its best group.  NS = 2 KT subsets of >= 4 points put the bound near the (1.1 KT)-th nearest seed
// point.  Not inlined: runs once per query.
template <int KT, int SLOTF4 = kBlockF4>
__device__ __noinline__ float seed_bound(const float4* ring4, int nseed, float a0, float a1, float a2, float qq,
                                         float E) {
  constexpr int NS = KT == 1 ? 1 : (KT == 4 ? 16 : 2 * KT);
  constexpr int UG = NS >= 4 ? NS / 4 : 1;  // groups per unrolled step: subset index stays static
  static_assert(UG <= kBlockGroups, "a step stays inside one block");
  const float INF = __int_as_float(0x7f800000);
  float mins[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) mins[i] = INF;
  const float2 A0 = make_float2(a0, a0), A1 = make_float2(a1, a1), A2 = make_float2(a2, a2);
  if (KT == 1) {
    // K = 1: the group of four seed points that holds the smallest expanded-form value, then the EXACT
    // distances of that one group: the distance of a real point bounds the nearest distance without an E
    // floor under it (a self-search starts at U = 0, a chamfer pair of near-identical clouds at its true, tiny
    // nearest distance), for 16 instead of 45 instructions per seed group (the all-exact seed was 27 % of the
    // K = 1 search's instructions on the chamfer shape)
    float best = INF;
    int bg = 0;
    for (int s = 0; s < nseed; ++s) {
      const float4* tp = ring4 + s * SLOTF4;
#pragma unroll 4
      for (int g = 0; g < kBlockGroups; ++g) {
        const float4 X = tp[g], Y = tp[kBlockGroups + g], Z = tp[2 * kBlockGroups + g], W = tp[3 * kBlockGroups + g];
        float2 s01 = make_float2(W.x, W.y), s23 = make_float2(W.z, W.w);
        s01 = __ffma2_rn(A0, make_float2(X.x, X.y), s01);
        s23 = __ffma2_rn(A0, make_float2(X.z, X.w), s23);
        s01 = __ffma2_rn(A1, make_float2(Y.x, Y.y), s01);
        s23 = __ffma2_rn(A1, make_float2(Y.z, Y.w), s23);
        s01 = __ffma2_rn(A2, make_float2(Z.x, Z.y), s01);
        s23 = __ffma2_rn(A2, make_float2(Z.z, Z.w), s23);
        const float m = fminf(fminf(s01.x, s01.y), fminf(s23.x, s23.y));
        const bool better = m < best;
        best = better ? m : best;
        bg = better ? s * SLOTF4 + g : bg;
      }
    }
    if (!(best < INF)) return INF;  // no valid seed point
    const float4* tg = ring4 + bg;
    const float4 W = tg[3 * kBlockGroups];
    float d4[4];
    exact4(-0.5f * a0, -0.5f * a1, -0.5f * a2, tg[0], tg[kBlockGroups], tg[2 * kBlockGroups], d4);
    const float w4[4] = {W.x, W.y, W.z, W.w};
    float m = INF;
#pragma unroll
    for (int i = 0; i < 4; ++i) m = fminf(m, (w4[i] == INF) ? INF : d4[i]);
    return m;
  }
  for (int s = 0; s < nseed; ++s) {
    const float4* tp = ring4 + s * SLOTF4;  // SLOTF4: float4 between consecutive ring slots
#pragma unroll 1
    for (int g0 = 0; g0 < kBlockGroups; g0 += UG) {
#pragma unroll
      for (int u = 0; u < UG; ++u) {
        const int g = g0 + u;
        const float4 X = tp[g], Y = tp[kBlockGroups + g], Z = tp[2 * kBlockGroups + g], W = tp[3 * kBlockGroups + g];
        float2 s01 = make_float2(W.x, W.y), s23 = make_float2(W.z, W.w);
        s01 = __ffma2_rn(A0, make_float2(X.x, X.y), s01);
        s23 = __ffma2_rn(A0, make_float2(X.z, X.w), s23);
        s01 = __ffma2_rn(A1, make_float2(Y.x, Y.y), s01);
        s23 = __ffma2_rn(A1, make_float2(Y.z, Y.w), s23);
        s01 = __ffma2_rn(A2, make_float2(Z.x, Z.y), s01);
        s23 = __ffma2_rn(A2, make_float2(Z.z, Z.w), s23);
        const float s4[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float& m = mins[(u * 4 + i) % NS];
          m = fminf(m, s4[i]);
        }
      }
    }
  }
  auto up = [&](float us) { return us < INF ? __fadd_rn(__fadd_rn(us, qq), E) : INF; };
  if (NS == 1) return up(mins[0]);
  if (NS != 2 * KT) {  // KT = 4: 4th smallest of 16
    sort_floats<NS, 0, NS>(mins);
    return up(mins[KT - 1]);
  }
  // KT-th smallest of 2 KT values: sort both halves, then max_i min(A[i], B[KT-1-i])
  constexpr int H = NS / 2;
  sort_floats<H, 0, NS>(mins);
  sort_floats<H, H, NS>(mins);
  float U = fminf(mins[0], mins[H + H - 1]);
#pragma unroll
  for (int i = 1; i < H; ++i) U = fmaxf(U, fminf(mins[i], mins[H + H - 1 - i]));
  return up(U);
}

}  // namespace
}  // namespace pops
