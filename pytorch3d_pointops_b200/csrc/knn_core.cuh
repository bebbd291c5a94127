// Pieces shared by the KNN scan kernels (knn.cu: tiled brute force; knn_prune.cu: Hilbert-ordered,
// box-pruned D = 3 search): constants, register sorting networks on 64-bit keys, and the
// warp-converged candidate flush.
#pragma once
#include "common.cuh"
#include "knn_order.cuh"

namespace pops {

constexpr int kGroup = 4;             // points per filter group (one float4 per SoA row)
constexpr int kChunk = 4;             // groups between candidate-buffer overflow checks
constexpr int kPadPoints = kGroup * kChunk;  // SoA rows padded to a multiple of this
constexpr int kSurvCap = 16;           // survivor keys a lane may hold between two merges
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

constexpr int kBufCap = 16;  // candidate buffer capacity (groups) per query


// ---- register sorting networks on 64-bit keys ------------------------------------------------
__device__ __forceinline__ void ce64(uint64_t& lo, uint64_t& hi) {  // lo <- min, hi <- max
  const bool sw = hi < lo;
  const uint64_t a = sw ? hi : lo, b = sw ? lo : hi;
  lo = a;
  hi = b;
}

// Batcher odd-even merge sort of 16 keys: 63 compare-exchanges (pairs generated offline and
// verified with the 0-1 principle); constexpr tables so that every index resolves statically.
constexpr int kSort16N = 63;
__device__ constexpr unsigned char kSort16A[kSort16N] = {0,2,4,6,8,10,12,14,0,1,4,5,8,9,12,13,1,5,9,13,0,1,2,3,8,9,10,11,2,3,10,11,1,3,5,9,11,13,0,1,2,3,4,5,6,7,4,5,6,7,2,3,6,7,10,11,1,3,5,7,9,11,13};
__device__ constexpr unsigned char kSort16B[kSort16N] = {1,3,5,7,9,11,13,15,2,3,6,7,10,11,14,15,2,6,10,14,4,5,6,7,12,13,14,15,4,5,12,13,2,4,6,10,12,14,8,9,10,11,12,13,14,15,8,9,10,11,4,5,8,9,12,13,2,4,6,8,10,12,14};
__device__ __forceinline__ void sort16(uint64_t (&v)[16]) {
#pragma unroll
  for (int e = 0; e < kSort16N; ++e) ce64(v[kSort16A[e]], v[kSort16B[e]]);
}

// v is bitonic -> ascending
template <int N>
__device__ __forceinline__ void bitonic_merge(uint64_t (&v)[N]) {
#pragma unroll
  for (int k = N / 2; k >= 1; k /= 2) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if ((i & k) == 0) ce64(v[i], v[i + k]);
    }
  }
}

// sorted insertion of one key into the ascending register list (branch-free, all compares
// against the OLD list, so the KT steps are independent)
template <int KT>
__device__ __forceinline__ void insert_network(uint64_t (&Lr)[KT], uint64_t key) {
  bool lt[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) lt[k] = key < Lr[k];
#pragma unroll
  for (int k = KT - 1; k >= 1; --k) Lr[k] = lt[k - 1] ? Lr[k - 1] : (lt[k] ? key : Lr[k]);
  Lr[0] = lt[0] ? key : Lr[0];
}

static_assert(kSurvCap == 16, "sort16 assumes 16 survivor slots");

// Batcher odd-even merge sort of v[OFF .. OFF+N), N in {4, 8, 16}, ascending, from compare-exchange tables
// (generated offline, verified with the 0-1 principle): one flat, fully unrolled loop, so every index is
// static and the values stay in registers.  (The nested-loop formulation of the same network left one
// stage rolled with run-time indices: the whole array then lives in local memory.)
constexpr int kSort4N = 5;
__device__ constexpr unsigned char kSort4A[kSort4N] = {0,2,0,1,1};
__device__ constexpr unsigned char kSort4B[kSort4N] = {1,3,2,3,2};
constexpr int kSort8N = 19;
__device__ constexpr unsigned char kSort8A[kSort8N] = {0,2,4,6,0,1,4,5,1,5,0,1,2,3,2,3,1,3,5};
__device__ constexpr unsigned char kSort8B[kSort8N] = {1,3,5,7,2,3,6,7,2,6,4,5,6,7,4,5,2,4,6};
template <int N, int OFF, int TOTAL>
__device__ __forceinline__ void sort_floats(float (&v)[TOTAL]) {
  if constexpr (N == 4 || N == 8 || N == 16) {
    constexpr int NCE = N == 4 ? kSort4N : (N == 8 ? kSort8N : kSort16N);
#pragma unroll
    for (int e = 0; e < NCE; ++e) {
      const int ia = OFF + (N == 4 ? kSort4A[e] : (N == 8 ? kSort8A[e] : kSort16A[e]));
      const int ib = OFF + (N == 4 ? kSort4B[e] : (N == 8 ? kSort8B[e] : kSort16B[e]));
      const float lo = fminf(v[ia], v[ib]);
      const float hi = fmaxf(v[ia], v[ib]);
      v[ia] = lo;
      v[ib] = hi;
    }
  } else {  // any other power of two (K = 32 seeds sort 2 x 32 values): the same network from nested loops
#pragma unroll
    for (int p = 1; p < N; p *= 2) {
#pragma unroll
      for (int k = p; k >= 1; k /= 2) {
#pragma unroll
        for (int j = k % p; j <= N - 1 - k; j += 2 * k) {
#pragma unroll
          for (int i = 0; i < k; ++i) {
            if (i <= N - j - k - 1 && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
              const float lo = fminf(v[OFF + i + j], v[OFF + i + j + k]);
              const float hi = fmaxf(v[OFF + i + j], v[OFF + i + j + k]);
              v[OFF + i + j] = lo;
              v[OFF + i + j + k] = hi;
            }
          }
        }
      }
    }
  }
}


// Merge one lane's `ns` survivor keys (column S, stride SSTRIDE) into its ascending K-list kept in
// the OUTPUT arrays (od, oi); warp-converged, `ns_max` = the largest ns in the warp.  Few
// survivors -> branch-free insertion network per survivor; many -> sort network + bitonic merge.
// `fresh`: the lane's list is empty and its row not yet written -- nothing is read.
// Returns min(dkt, the list's K-th distance) for lanes that merged, dkt otherwise.
template <int KT, int SSTRIDE>
__device__ __forceinline__ float knn_merge_global(const uint64_t* S, int ns, int ns_max, int K, float* od,
                                                  int64_t* oi, float dkt, bool fresh = false) {
  constexpr int KR = KT;
  uint64_t Lr[KR];
  const bool mine = ns > 0;  // only lanes that hold survivors touch their list
  // a fresh list reads as (+inf, 0xFFFFFFFF) in its K slots, the value an empty row used to be filled with
#pragma unroll
  for (int k = 0; k < KR; ++k) Lr[k] = (fresh && k < K) ? 0x7F800000FFFFFFFFull : kEmptyKey;
  if (mine && !fresh) {
    if (KR >= 4 && (K & 3) == 0) {  // rows are 16-byte aligned: 128-bit loads
#pragma unroll
      for (int k4 = 0; k4 < KR / 4; ++k4) {
        if (k4 * 4 < K) {
          const float4 dv = reinterpret_cast<const float4*>(od)[k4];
          const longlong2 i01 = reinterpret_cast<const longlong2*>(oi)[k4 * 2];
          const longlong2 i23 = reinterpret_cast<const longlong2*>(oi)[k4 * 2 + 1];
          Lr[k4 * 4 + 0] = make_key(dv.x, static_cast<uint32_t>(i01.x));
          Lr[k4 * 4 + 1] = make_key(dv.y, static_cast<uint32_t>(i01.y));
          Lr[k4 * 4 + 2] = make_key(dv.z, static_cast<uint32_t>(i23.x));
          Lr[k4 * 4 + 3] = make_key(dv.w, static_cast<uint32_t>(i23.y));
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KR; ++k)
        if (k < K) Lr[k] = make_key(od[k], static_cast<uint32_t>(oi[k]));
    }
  }
  if (ns_max <= 5 || KR < 4) {
    for (int s2 = 0; s2 < ns_max; ++s2) {
      const uint64_t key = (s2 < ns) ? S[s2 * SSTRIDE] : kEmptyKey;
      insert_network<KR>(Lr, key);
    }
  } else {
    uint64_t Sr[kSurvCap];
#pragma unroll
    for (int s2 = 0; s2 < kSurvCap; ++s2) Sr[s2] = (s2 < ns) ? S[s2 * SSTRIDE] : kEmptyKey;
    sort16(Sr);
    if (KR <= 8 && __all_sync(0xffffffffu, fresh || ns == 0)) {
      // every list that receives survivors is still empty (a warp's first flush): the sorted survivors ARE the
      // list, no merge network.  (K <= 8 only: 326 -> 320 us on the T shape at K = 8; for K = 16 / 32 the second
      // code path costs more than the skipped network saves, 576 -> 600 us.)
#pragma unroll
      for (int i = 0; i < KR; ++i) Lr[i] = (Sr[i] < Lr[i]) ? Sr[i] : Lr[i];  // keeps the (+inf, none) filler where Sr is empty
    } else {
      // K smallest of (Lr U Sr): C[i] = min(Lr[i], Sr[KR-1-i]) is bitonic, then merge
#pragma unroll
      for (int i = 0; i < KR; ++i) {
        const int si = KR - 1 - i;
        if (si < kSurvCap) Lr[i] = (Sr[si] < Lr[i]) ? Sr[si] : Lr[i];
      }
      bitonic_merge<KR>(Lr);
    }
  }
  if (mine) {
    if (KR >= 4 && (K & 3) == 0) {
#pragma unroll
      for (int k4 = 0; k4 < KR / 4; ++k4) {
        if (k4 * 4 < K) {
          reinterpret_cast<float4*>(od)[k4] =
              make_float4(key_dist(Lr[k4 * 4]), key_dist(Lr[k4 * 4 + 1]), key_dist(Lr[k4 * 4 + 2]),
                          key_dist(Lr[k4 * 4 + 3]));
          reinterpret_cast<longlong2*>(oi)[k4 * 2] =
              make_longlong2(static_cast<long long>(Lr[k4 * 4] & 0xFFFFFFFFull),
                             static_cast<long long>(Lr[k4 * 4 + 1] & 0xFFFFFFFFull));
          reinterpret_cast<longlong2*>(oi)[k4 * 2 + 1] =
              make_longlong2(static_cast<long long>(Lr[k4 * 4 + 2] & 0xFFFFFFFFull),
                             static_cast<long long>(Lr[k4 * 4 + 3] & 0xFFFFFFFFull));
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        if (k < K) {
          od[k] = key_dist(Lr[k]);
          oi[k] = static_cast<int64_t>(Lr[k] & 0xFFFFFFFFull);
        }
      }
    }
    // K == KT (the usual K = 1, 4, 8, 16, 32) reads the last register; otherwise the compiler indexes the
    // list through local memory (KT stores and a load), which then stay on this cold path
    uint64_t worst = Lr[KR - 1];
    if (K != KR) {
      worst = kEmptyKey;
#pragma unroll
      for (int k = 0; k < KR; ++k)
        if (k == K - 1) worst = Lr[k];
    }
    // the list's K-th entry (+inf while it is not full), or a tighter bound the caller had
    dkt = fminf(dkt, key_dist(worst));
  }
  return dkt;
}

// ---------------------------------------------------------------------------------------------
// flush: drain ONE query's candidate buffer.  A real (non-inlined) function: it is big, runs
// rarely, and must exist once, not once per call site and query slot -- the scan loop has to stay
// resident in the instruction cache.  Called warp-converged and kept converged inside: rare events
// never sit inside a dense loop.
//   fill   per buffered group, all 4 points get the exact unfused distance (branch-free, packed
//          f32x2 sub/mul, scalar adds -- IEEE, never fused); points with d <= dk are appended as
//          64-bit keys to the lane's survivor column (predicated);
//   merge  KT > 0: the list (kept in the OUTPUT arrays) is pulled into registers by the lanes that
//          hold survivors; few survivors -> branch-free insertion network per survivor, many ->
//          sort network + bitonic merge.  KT == 0 (any K): sorted survivors are merged backward in
//          place into the shared-memory list.
// Returns the query's new K-th distance (+inf while the list is not full).
template <int DT, int NORM, bool EXP, int THREADS, int KT, int RS, int QPB>
__device__ __noinline__ float knn_flush_one(const float* tile, const unsigned short* cand_col, int c_end,
                                            uint64_t* L, uint64_t* S, float4 qv, float dkt, int j0,
                                            int L2, int K, float* od, int64_t* oi) {
  if (!__any_sync(0xffffffffu, c_end > 0)) return dkt;
  const float qarr[4] = {qv.x, qv.y, qv.z, qv.w};
  float q[DT];
#pragma unroll
  for (int d = 0; d < DT; ++d) q[d] = qarr[d];
  const float INF = __int_as_float(0x7f800000);
    int c = 0;
    for (;;) {
      if (!__any_sync(0xffffffffu, c < c_end)) break;
      // ---- fill ----
      int ns = 0;
      while (c < c_end && ns <= kSurvCap - kGroup) {
        const int g = cand_col[c * QPB];
        ++c;
        float4 X[DT];
#pragma unroll
        for (int d = 0; d < DT; ++d) X[d] = reinterpret_cast<const float4*>(tile + d * RS)[g];
        float dist[kGroup];
        if (NORM == 2) {
          float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < DT; ++d) {
            const float2 qd = make_float2(q[d], q[d]);
            const float2 d01 = __fadd2_rn(qd, make_float2(-X[d].x, -X[d].y));
            const float2 d23 = __fadd2_rn(qd, make_float2(-X[d].z, -X[d].w));
            const float2 t01 = __fmul2_rn(d01, d01), t23 = __fmul2_rn(d23, d23);
            // scalar adds on purpose: ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2
            // (even with -fmad=false), which would break bit parity with the unfused reference
            acc01 = d == 0 ? t01 : make_float2(__fadd_rn(acc01.x, t01.x), __fadd_rn(acc01.y, t01.y));
            acc23 = d == 0 ? t23 : make_float2(__fadd_rn(acc23.x, t23.x), __fadd_rn(acc23.y, t23.y));
          }
          dist[0] = acc01.x; dist[1] = acc01.y; dist[2] = acc23.x; dist[3] = acc23.y;
        } else {
#pragma unroll
          for (int i = 0; i < kGroup; ++i) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < DT; ++d) {
              const float xv = i == 0 ? X[d].x : (i == 1 ? X[d].y : (i == 2 ? X[d].z : X[d].w));
              const float term = dist_term<NORM>(q[d], xv);
              acc = (d == 0) ? term : __fadd_rn(acc, term);
            }
            dist[i] = acc;
          }
        }
        const int jg = j0 + g * kGroup;
#pragma unroll
        for (int i = 0; i < kGroup; ++i) {
          if (dist[i] <= dkt && jg + i < L2) {
            S[ns * THREADS] = make_key(dist[i], static_cast<uint32_t>(jg + i));
            ++ns;
          }
        }
      }
      __syncwarp();
      // ---- merge ----
      const int ns_max = __reduce_max_sync(0xffffffffu, ns);
      if (ns_max > 0) {
        if (KT > 0) {
          dkt = knn_merge_global<(KT > 0 ? KT : 1), THREADS>(S, ns, ns_max, K, od, oi, dkt);
        } else if (ns > 0) {
          for (int a2 = 1; a2 < ns; ++a2) {  // insertion sort of the survivors
            const uint64_t key = S[a2 * THREADS];
            int b2 = a2 - 1;
            while (b2 >= 0) {
              const uint64_t prev = S[b2 * THREADS];
              if (prev <= key) break;
              S[(b2 + 1) * THREADS] = prev;
              --b2;
            }
            S[(b2 + 1) * THREADS] = key;
          }
          int r = 0;  // survivors that belong to the K smallest of (list U survivors)
          while (r < ns && r < K && S[r * THREADS] < L[static_cast<size_t>(K - 1 - r) * QPB]) ++r;
          int i = K - 1 - r, jj = r - 1, o = K - 1;  // backward in-place merge
          while (jj >= 0) {
            const uint64_t sv = S[jj * THREADS];
            uint64_t lv = 0;
            if (i >= 0) lv = L[static_cast<size_t>(i) * QPB];
            if (i >= 0 && lv > sv) {
              L[static_cast<size_t>(o) * QPB] = lv;
              --i;
            } else {
              L[static_cast<size_t>(o) * QPB] = sv;
              --jj;
            }
            --o;
          }
          const uint64_t worst = L[static_cast<size_t>(K - 1) * QPB];
          if (worst != kEmptyKey) dkt = key_dist(worst);
        }
      }
      __syncwarp();
    }
    return dkt;
}



// knn_prune.cu: D = 3, L2, K <= 32 search over the pre-pass output (curve order, blocks, boxes).
int knn_prune_search(const KnnOrderBuffers& ob, const int64_t* len1, const int64_t* len2, int N, int P1,
                     int P2, int K, int64_t* idx, float* dists, cudaStream_t st);

// knn_tc.cu: 32 <= D <= 256, L2, K <= 16 on the tensor cores (tcgen05 filter + exact re-rank).
bool knn_tc_supported(int64_t P1, int64_t P2, int64_t D, int64_t K, int norm);
size_t knn_tc_workspace_bytes(int64_t N, int64_t P1, int64_t P2);
int knn_tc_search(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1,
                  int P2, int D, int K, int64_t* idx, float* dists, void* ws, unsigned char** flags_out,
                  const unsigned** flag_count_out, unsigned* flag_limit_out, cudaStream_t st);

}  // namespace pops
