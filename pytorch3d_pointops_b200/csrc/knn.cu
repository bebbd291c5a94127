// K nearest neighbours, forward + backward, for sm_100a.
//
// Replaces csrc/knn/{knn.cu,knn_cpu.cpp} of the reference AND the sort/gather post-pass of
// functions/knn.py:77-89.  Contract = the reference CPU path (knn_cpu.cpp:13-69): the K
// lexicographically smallest (dist, idx) per query, ascending, dist computed as the unfused
// float32 sum  fl(fl(dx*dx) + fl(dy*dy)) + ...  (or |dx| + |dy| + ... for norm 1).
//
// Design (DESIGN.md "KNN"):
//   pack pass   p2 (N,P2,D) AoS -> per-cloud SoA rows x[],y[],z[](,w=|p|^2) padded with
//               sentinels, plus max|coord| per cloud (bounds the filter's rounding error).
//   scan pass   one CTA = QPB queries of one cloud.  p2 streams through shared memory in
//               tiles moved by TMA bulk copies (cp.async.bulk + mbarrier, 2 stages).  Each thread
//               owns Q queries and, per group of 4 points, evaluates a cheap FILTER:
//                 D==3, L2:  s = w + (-2qx)x + (-2qy)y + (-2qz)z   (3 FFMA, issued as packed
//                            FFMA2 over point pairs) tested against T = (dk - |q|^2) + E,
//                            E >= the filter's worst-case rounding error, so no true
//                            neighbour is ever dropped;
//                 otherwise: the exact distance itself, tested against dk.
//               Groups that pass are appended (branch-free, predicated) to a small per-query
//               candidate buffer in shared memory.  Buffers are drained in warp-converged
//               FLUSH phases: exact unfused distance, 64-bit key (dist_bits<<32 | idx), sorted
//               insertion into the per-query top-K list (shared memory, column per query).
//               Only the exact key decides membership and order -> bit-exact vs the oracle.
//   generic     any D / any K fallback: thread per query, exact distance, list in shared or
//               global memory.
#include <cfloat>
#include <cstdlib>

#include "knn_core.cuh"

namespace pops {

// ---------------------------------------------------------------------------------------------
// pack pass
// ---------------------------------------------------------------------------------------------
// soa layout: [n][row][P2pad], rows = DT (+1 for w when EXP).  j >= len2: sentinel
// (EXP: x=y=z=0, w=+inf -> s=+inf;  else: row0=+inf -> d=+inf), never a candidate by itself.
template <int DT, bool EXP>
__global__ void knn_pack_kernel(const float* __restrict__ p2, const int64_t* __restrict__ len2,
                                int P2, int P2pad, float* __restrict__ soa,
                                unsigned* __restrict__ maxabs_bits) {
  constexpr int ROWS = DT + (EXP ? 1 : 0);
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int64_t L = len2[n];
  L = L < 0 ? 0 : (L > P2 ? P2 : L);
  unsigned m = 0u;  // max |coordinate| as a bit pattern: +inf and NaN rank above every finite value
  if (j < P2pad) {
    float v[DT];
    float w = 0.0f;
    if (j < L) {
      const float* src = p2 + (static_cast<size_t>(n) * P2 + j) * DT;
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        v[d] = src[d];
        m = max(m, abs_bits(v[d]));
        w = fmaf(v[d], v[d], w);
      }
    } else {
#pragma unroll
      for (int d = 0; d < DT; ++d) v[d] = 0.0f;
      if (EXP) {
        w = __int_as_float(0x7f800000);
      } else {
        v[0] = __int_as_float(0x7f800000);
      }
    }
    float* dst = soa + static_cast<size_t>(n) * ROWS * P2pad + j;
#pragma unroll
    for (int d = 0; d < DT; ++d) dst[static_cast<size_t>(d) * P2pad] = v[d];
    if (EXP) dst[static_cast<size_t>(DT) * P2pad] = w;
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m > 0u) atomicMax(maxabs_bits + n, m);
}

// max |coord| over the valid rows of p (N,P,D) -> atomicMax into maxabs_bits[n]
__global__ void maxabs_kernel(const float* __restrict__ p, const int64_t* __restrict__ len, int P,
                              int D, unsigned* __restrict__ maxabs_bits) {
  const int n = blockIdx.y;
  int64_t L = len[n];
  L = L < 0 ? 0 : (L > P ? P : L);
  const size_t total = static_cast<size_t>(L) * D;
  const float* src = p + static_cast<size_t>(n) * P * D;
  unsigned m = 0u;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = max(m, abs_bits(src[i]));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m > 0u) atomicMax(maxabs_bits + n, m);
}

// ---------------------------------------------------------------------------------------------
// scan pass (tiled, buffered candidates)
// ---------------------------------------------------------------------------------------------
struct KnnScanParams {
  const float* p1;
  const float* soa;
  const int64_t* len1;
  const int64_t* len2;
  const unsigned* maxabs_bits;
  int64_t* idx;
  float* dists;
  int P1, P2, P2pad, K;
};

// Shared-memory carve-up.  RS = points per tile row (compile time, so tile loads in the hot loop
// are immediate-offset LDS.128).  ONE tile stage: the TMA refill of a CTA overlaps with the other
// resident CTA's compute; a single large tile halves the number of forced end-of-tile flushes.
template <int DT, bool EXP, bool GL, int Q, int THREADS, int RS>
struct KnnSmem {
  static constexpr int ROWS = DT + (EXP ? 1 : 0);
  static constexpr int QPB = Q * THREADS;
  static constexpr size_t tiles_off = 64;
  static constexpr size_t tiles_bytes = size_t(ROWS) * RS * 4 + 64;  // +64: the prefetch over-read
  static constexpr size_t lists_off = tiles_off + tiles_bytes;
  // GL: the top-K lists live in the OUTPUT arrays (global memory), not in shared memory
  static __host__ __device__ size_t lists_bytes(int K) { return GL ? 0 : size_t(K) * QPB * 8; }
  static __host__ __device__ size_t cand_off(int K) { return lists_off + lists_bytes(K); }
  static constexpr size_t cand_bytes = size_t(kBufCap) * QPB * 2;
  static __host__ __device__ size_t surv_off(int K) { return (cand_off(K) + cand_bytes + 7) / 8 * 8; }
  static constexpr size_t surv_bytes = size_t(kSurvCap) * THREADS * 8;
  static __host__ __device__ size_t total(int K) { return surv_off(K) + surv_bytes; }
};

template <int DT, int NORM, bool EXP, int Q, int THREADS, int KT, int RS>
__global__ void __launch_bounds__(THREADS, (THREADS > 192 ? 1 : ((KT > 0 && KT <= 16) ? 3 : 2)))
knn_scan_kernel(const KnnScanParams prm) {
  constexpr bool GL = KT > 0;  // register-merge variants keep their lists in the output arrays
  using SM = KnnSmem<DT, EXP, GL, Q, THREADS, RS>;
  constexpr int ROWS = SM::ROWS;
  constexpr int QPB = SM::QPB;
  extern __shared__ __align__(128) unsigned char smem[];

  const int n = blockIdx.y;
  const int q_base = blockIdx.x * QPB;
  const int tid = threadIdx.x;
  const int K = prm.K, P2pad = prm.P2pad;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));

  int64_t* out_idx = prm.idx + (static_cast<size_t>(n) * prm.P1) * K;
  float* out_d = prm.dists + (static_cast<size_t>(n) * prm.P1) * K;
  if (prm.maxabs_bits[n] >= kDirtyBits) return;  // non-finite / huge coordinates: the exact generic kernel answers this cloud

  // CTA entirely beyond lengths1[n] (or nothing to search): rows are (0, 0).
  if (q_base >= L1 || L2 == 0) {
    const int rows = min(QPB, prm.P1 - q_base);
    for (int e = tid; e < rows * K; e += THREADS) {
      const size_t row = static_cast<size_t>(q_base) + e / K;
      out_idx[row * K + e % K] = 0;
      out_d[row * K + e % K] = 0.0f;
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float* tiles = reinterpret_cast<float*>(smem + SM::tiles_off);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + SM::lists_off);
  unsigned short* cand = reinterpret_cast<unsigned short*>(smem + SM::cand_off(K));
  uint64_t* surv = reinterpret_cast<uint64_t*>(smem + SM::surv_off(K));

  const int L2pad = (L2 + kPadPoints - 1) / kPadPoints * kPadPoints;  // <= P2pad
  const int num_tiles = (L2pad + RS - 1) / RS;
  const float* soa_n = prm.soa + static_cast<size_t>(n) * ROWS * P2pad;

  auto issue_tile = [&](int tile) {
    const int j0 = tile * RS;
    const int pts = min(RS, L2pad - j0);
    const uint32_t bytes = static_cast<uint32_t>(pts) * 4u;
    mbar_arrive_expect_tx(&bars[0], bytes * ROWS);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      tma_bulk_g2s(tiles + static_cast<size_t>(r) * RS, soa_n + static_cast<size_t>(r) * P2pad + j0, bytes,
                   &bars[0]);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) issue_tile(0);

  // ---- per-thread query state ---------------------------------------------------------------
  const float M = __uint_as_float(prm.maxabs_bits[n]);
  // E >= 130.2 * 2^-24 * M^2 bounds |filter - reference| (DESIGN.md "filter error bound").
  const float E = fmaf(M * M, 1.52587890625e-05f /* 2^-16 */, 1e-37f);
  constexpr uint32_t CSTRIDE = QPB * 2;  // bytes between consecutive entries of one candidate buffer
  const float INF = __int_as_float(0x7f800000);
  float a[Q][DT];   // EXP: -2 q_d (FFMA2 takes it as a broadcast scalar operand);  else q_d
  float qq[Q];
  float T[Q];       // filter threshold
  float dk[Q];      // current K-th distance (+inf while the list is not full)
  uint32_t cw[Q];   // shared-memory byte address of the next free candidate slot
  const uint32_t cand_base = smem_u32(cand) + static_cast<uint32_t>(tid) * 2u;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const int slot = tid + t * THREADS;
    const int qi = q_base + slot;
    const bool valid = qi < L1;
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      const float q = valid ? prm.p1[(static_cast<size_t>(n) * prm.P1 + qi) * DT + d] : 0.0f;
      s = fmaf(q, q, s);
      a[t][d] = EXP ? -2.0f * q : q;
    }
    qq[t] = s;
    dk[t] = valid ? INF : -1.0f;
    T[t] = valid ? (EXP ? FLT_MAX : INF) : -INF;
    cw[t] = cand_base + static_cast<uint32_t>(t) * (THREADS * 2u);
    if (GL) {
      // empty list = (+inf, 0xFFFFFFFF) in every slot; rows beyond lengths1 are final zeros
      if (qi < prm.P1) {
        float* od = out_d + static_cast<size_t>(qi) * K;
        int64_t* oi = out_idx + static_cast<size_t>(qi) * K;
        for (int k = 0; k < K; ++k) {
          od[k] = valid ? INF : 0.0f;
          oi[k] = valid ? static_cast<int64_t>(0xFFFFFFFFll) : 0;
        }
      }
    } else {
      for (int k = 0; k < K; ++k) lists[static_cast<size_t>(k) * QPB + slot] = kEmptyKey;
    }
  }

  // ---- flush glue: drain query t's candidate buffer (knn_flush_one, knn_core.cuh) ---------------
  auto flush = [&](int t, const float* tile, int j0) {
    const int slot = tid + t * THREADS;
    const uint32_t base = cand_base + static_cast<uint32_t>(t) * (THREADS * 2u);
    const int c_end = static_cast<int>((cw[t] - base) / CSTRIDE);
    cw[t] = base;
    float4 qv;  // a = -2q exactly (power-of-two scaling), so q = -a/2 exactly
    qv.x = EXP ? -0.5f * a[t][0] : a[t][0];
    qv.y = EXP ? -0.5f * a[t][1 % DT] : a[t][1 % DT];
    qv.z = EXP ? -0.5f * a[t][2 % DT] : a[t][2 % DT];
    qv.w = EXP ? -0.5f * a[t][3 % DT] : a[t][3 % DT];
    const int qi = q_base + slot;
    const float dkt = knn_flush_one<DT, NORM, EXP, THREADS, KT, RS, QPB>(
        tile, cand + slot, c_end, GL ? nullptr : lists + slot, surv + tid, qv, dk[t], j0, L2, K,
        out_d + static_cast<size_t>(qi) * K, out_idx + static_cast<size_t>(qi) * K);
    dk[t] = dkt;
    if (dkt >= 0.0f && dkt < INF) T[t] = EXP ? __fadd_rn(__fsub_rn(dkt, qq[t]), E) : dkt;
  };

  // ---- main loop over p2 tiles ------------------------------------------------------------------
  const uint32_t cw_limit = cand_base + static_cast<uint32_t>(kBufCap - kChunk) * CSTRIDE;
  for (int tile_i = 0; tile_i < num_tiles; ++tile_i) {
    const int j0 = tile_i * RS;
    const int pts = min(RS, L2pad - j0);
    const int ngroups = pts / kGroup;  // multiple of kChunk
    const float* tile = tiles;
    mbar_wait(&bars[0], tile_i & 1);

    // Groups are read with immediate-offset LDS.128; the next group's rows are loaded while the
    // current group is evaluated (the last prefetch reads one group past the tile: in-bounds shared
    // memory, never used).  ONE flush call site (overflow of any lane, or end of tile).
    const float4* tp = reinterpret_cast<const float4*>(tile);
    constexpr int SROWS = ROWS;
    float4 Xc[SROWS];
#pragma unroll
    for (int r = 0; r < SROWS; ++r) Xc[r] = tp[r * (RS / 4)];
    for (int g = 0; g < ngroups; g += kChunk) {
#pragma unroll
      for (int c = 0; c < kChunk; ++c) {
        float4 Xn[SROWS];
#pragma unroll
        for (int r = 0; r < SROWS; ++r) Xn[r] = tp[r * (RS / 4) + g + c + 1];
        const unsigned short g16 = static_cast<unsigned short>(g + c);
#pragma unroll
        for (int t = 0; t < Q; ++t) {
          float m;
          if (EXP) {
            float2 s01 = make_float2(Xc[DT].x, Xc[DT].y), s23 = make_float2(Xc[DT].z, Xc[DT].w);
#pragma unroll
            for (int d = 0; d < DT; ++d) {
              const float2 ad = make_float2(a[t][d], a[t][d]);
              s01 = __ffma2_rn(ad, make_float2(Xc[d].x, Xc[d].y), s01);
              s23 = __ffma2_rn(ad, make_float2(Xc[d].z, Xc[d].w), s23);
            }
            m = fminf(fminf(s01.x, s01.y), fminf(s23.x, s23.y));
          } else {
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int d = 0; d < DT; ++d) {
              const float qd = a[t][d];
              const float t0 = dist_term<NORM>(qd, Xc[d].x), t1 = dist_term<NORM>(qd, Xc[d].y);
              const float t2 = dist_term<NORM>(qd, Xc[d].z), t3 = dist_term<NORM>(qd, Xc[d].w);
              d0 = d == 0 ? t0 : __fadd_rn(d0, t0);
              d1 = d == 0 ? t1 : __fadd_rn(d1, t1);
              d2 = d == 0 ? t2 : __fadd_rn(d2, t2);
              d3 = d == 0 ? t3 : __fadd_rn(d3, t3);
            }
            m = fminf(fminf(d0, d1), fminf(d2, d3));
          }
          if (m <= T[t]) {  // predicated: one STS.U16 + one IADD
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(cw[t]), "h"(g16) : "memory");
            cw[t] += CSTRIDE;
          }
        }
#pragma unroll
        for (int r = 0; r < SROWS; ++r) Xc[r] = Xn[r];
      }
      // warp-converged: flush everything when ANY lane's buffer is nearly full, or at tile end
      uint32_t mx = cw[0];
#pragma unroll
      for (int t = 1; t < Q; ++t) mx = max(mx, cw[t] - static_cast<uint32_t>(t) * (THREADS * 2u));
      if (g + kChunk >= ngroups || __any_sync(0xffffffffu, mx > cw_limit)) {
#pragma unroll
        for (int t = 0; t < Q; ++t) flush(t, tile, j0);
      }
    }

    __syncthreads();  // everyone is done reading the tile
    if (tid == 0 && tile_i + 1 < num_tiles) {
      fence_proxy_async();
      issue_tile(tile_i + 1);
    }
  }

  // ---- write out ---------------------------------------------------------------------------------
  if (GL) {
    // the lists ARE the outputs; only slots beyond lengths2 (K > lengths2) still hold the empty
    // marker and become the reference's (0, 0) padding
    if (L2 < K) {
#pragma unroll
      for (int t = 0; t < Q; ++t) {
        const int qi = q_base + tid + t * THREADS;
        if (qi >= L1) continue;
        float* od = out_d + static_cast<size_t>(qi) * K;
        int64_t* oi = out_idx + static_cast<size_t>(qi) * K;
        for (int k = L2; k < K; ++k) {
          od[k] = 0.0f;
          oi[k] = 0;
        }
      }
    }
    return;
  }
  // sorted keys -> (idx, dist); empty slots and rows >= L1 are (0, 0)
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const int slot = tid + t * THREADS;
    const int qi = q_base + slot;
    if (qi >= prm.P1) continue;
    int64_t* oi = out_idx + static_cast<size_t>(qi) * K;
    float* od = out_d + static_cast<size_t>(qi) * K;
    for (int k = 0; k < K; ++k) {
      const uint64_t key = lists[static_cast<size_t>(k) * QPB + slot];
      const bool ok = key != kEmptyKey;
      oi[k] = ok ? static_cast<int64_t>(key & 0xFFFFFFFFull) : 0;
      od[k] = ok ? key_dist(key) : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// generic fallback: any D, any K.  Thread per query; p2 tile (AoS) and the CTA's queries in
// shared memory; exact distance; list column in shared memory when it fits, else in the
// global workspace.
// ---------------------------------------------------------------------------------------------
struct KnnGenericParams {
  const float* p1;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  int64_t* idx;
  float* dists;
  uint64_t* glists;  // [N][ceil(P1/THREADS)][K][THREADS] when lists live in global memory
  int P1, P2, D, K, TP;
  int lists_in_smem;
  const unsigned char* flags;  // [N][P1] or nullptr: when set, only flagged queries are (re)computed
  const unsigned* flag_count;  // with flags: number of flagged queries; this kernel runs only above
  unsigned flag_limit;         //   flag_limit of them (below, knn_exact_rows_kernel has done the work)
  const unsigned* dirty_bits;  // [N] or nullptr: when set, only clouds with dirty_bits[n] >= kDirtyBits
                               //   (non-finite or huge coordinates, skipped by the filtered search) are computed
};

template <int NORM, int THREADS>
__global__ void __launch_bounds__(THREADS)
knn_generic_kernel(const KnnGenericParams prm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = blockIdx.y, tid = threadIdx.x, q_base = blockIdx.x * THREADS;
  const int D = prm.D, K = prm.K, TP = prm.TP;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  const int qi = q_base + tid;
  const bool valid = qi < L1;
  bool wanted = true;
  if (prm.dirty_bits && prm.dirty_bits[n] < kDirtyBits) return;
  if (prm.flags) {  // exact recomputation of the queries the tensor-core path could not certify
    if (*prm.flag_count <= prm.flag_limit) return;
    wanted = qi < prm.P1 && prm.flags[static_cast<size_t>(n) * prm.P1 + qi] != 0;
    if (!__syncthreads_or(wanted ? 1 : 0)) return;
  }

  float* qsm = reinterpret_cast<float*>(smem);                   // [D][THREADS] (transposed)
  float* tile = qsm + static_cast<size_t>(D) * THREADS;          // [TP][D]
  uint64_t* lists = prm.lists_in_smem
                        ? reinterpret_cast<uint64_t*>(smem + align_up((size_t(D) * THREADS + size_t(TP) * D) * 4, 8))
                        : prm.glists + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * K * THREADS;
  for (int d = 0; d < D; ++d)
    qsm[d * THREADS + tid] = valid ? prm.p1[(static_cast<size_t>(n) * prm.P1 + qi) * D + d] : 0.0f;
  for (int k = 0; k < K; ++k) lists[static_cast<size_t>(k) * THREADS + tid] = kEmptyKey;
  uint64_t worst = kEmptyKey;

  const float* p2n = prm.p2 + static_cast<size_t>(n) * prm.P2 * D;
  for (int j0 = 0; j0 < L2; j0 += TP) {
    const int pts = min(TP, L2 - j0);
    __syncthreads();
    for (int e = tid; e < pts * D; e += THREADS) tile[e] = p2n[static_cast<size_t>(j0) * D + e];
    __syncthreads();
    if (!valid) continue;
    for (int jl = 0; jl < pts; ++jl) {
      float d = 0.0f;
      const float* pt = tile + jl * D;
      for (int dd = 0; dd < D; ++dd) {
        const float term = dist_term<NORM>(qsm[dd * THREADS + tid], pt[dd]);
        d = __fadd_rn(d, term);
      }
      // total order on (dist, idx): finite < +inf < NaN, ties by lower index (common.cuh)
      const uint64_t key = make_key_total(d, static_cast<uint32_t>(j0 + jl));
      if (key < worst) {
        int k = K - 1;
        while (k > 0) {
          const uint64_t prev = lists[static_cast<size_t>(k - 1) * THREADS + tid];
          if (prev <= key) break;
          lists[static_cast<size_t>(k) * THREADS + tid] = prev;
          --k;
        }
        lists[static_cast<size_t>(k) * THREADS + tid] = key;
        worst = lists[static_cast<size_t>(K - 1) * THREADS + tid];
      }
    }
  }
  if (qi < prm.P1 && wanted) {
    int64_t* oi = prm.idx + (static_cast<size_t>(n) * prm.P1 + qi) * K;
    float* od = prm.dists + (static_cast<size_t>(n) * prm.P1 + qi) * K;
    for (int k = 0; k < K; ++k) {
      const uint64_t key = lists[static_cast<size_t>(k) * THREADS + tid];
      const bool ok = valid && key != kEmptyKey;
      oi[k] = ok ? static_cast<int64_t>(key & 0xFFFFFFFFull) : 0;
      od[k] = ok ? key_dist(key) : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward (knn_cpu.cpp:75-128)
// ---------------------------------------------------------------------------------------------
// One thread per (n, i1, d): accumulates grad_p1 in a register over k (no atomics on p1),
// scatters -diff into grad_p2 with red.global.add.f32.
template <int NORM>
__global__ void knn_backward_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                    const int64_t* __restrict__ len1,
                                    const int64_t* __restrict__ len2,
                                    const int64_t* __restrict__ idx,
                                    const float* __restrict__ grad_dists, int N, int P1, int P2,
                                    int D, int K, float* __restrict__ grad_p1,
                                    float* __restrict__ grad_p2) {
  const size_t total = static_cast<size_t>(N) * P1 * D;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(e % D);
    const size_t row = e / D;  // n*P1 + i1
    const int n = static_cast<int>(row / P1);
    const int i1 = static_cast<int>(row % P1);
    float acc = 0.0f;
    int64_t L1 = len1[n], L2 = len2[n];
    if (i1 < L1) {
      const int kmax = static_cast<int>(L2 < K ? (L2 < 0 ? 0 : L2) : K);
      const float a = p1[e];
      for (int k = 0; k < kmax; ++k) {
        const int64_t i2 = idx[row * K + k];
        if (i2 < 0 || i2 >= P2) continue;  // -1 = padding (ball query)
        const float g = grad_dists[row * K + k];
        const float b = p2[(static_cast<size_t>(n) * P2 + i2) * D + d];
        float diff;
        // same float ops, same order as knn_cpu.cpp:113-122; intrinsics forbid contraction
        if (NORM == 1) {
          diff = __fmul_rn(g, (a > b) ? 1.0f : -1.0f);
        } else {
          diff = __fmul_rn(__fmul_rn(2.0f, g), __fsub_rn(a, b));
        }
        acc = __fadd_rn(acc, diff);
        atomicAdd(grad_p2 + (static_cast<size_t>(n) * P2 + i2) * D + d, -diff);
      }
    }
    grad_p1[e] = acc;
  }
}

// D <= 4, K <= 1024: one thread per (row, k) ENTRY.  The CTA owns whole rows (rows_per_cta * K entries):
// idx / grad_dists are read once, perfectly coalesced; the entry's diff vector goes to shared memory
// and to grad_p2 -- for D = 3 as ONE 16-byte vector reduction (red.global.add.v4.f32) into a
// (N*P2) float4 scratch that knn_backward_compact_kernel folds into the 12-byte rows afterwards: a
// third of the L2 atomic operations, which are what bounds this kernel.  Then one thread per
// (row, d) sums the row's K diffs in k order: grad_p1 is bit-exact vs knn_cpu.cpp:104-124.
// Measured on the T shape (8.4 M entries, 106 us + 8.5 us compaction) and dropped: the p2 cloud staged in
// shared memory, so that idx -> p2[idx] is no random fetch (122 us: the fetches are not the limit, the
// reduction rate of the L2 slices is); a cloud's grad_p2 accumulated in shared memory and added to global
// memory once per CTA (222 us: a float atomicAdd on shared memory is a compare-and-swap loop,
// ATOMS.CAST.SPIN, three per entry).
__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float2* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

template <int NORM, int DT>
__global__ void __launch_bounds__(256)
knn_backward_rows_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                         const int64_t* __restrict__ len1, const int64_t* __restrict__ len2,
                         const int64_t* __restrict__ idx, const float* __restrict__ grad_dists,
                         int64_t total_rows, int P1, int P2, int K, int rows_per_cta,
                         float* __restrict__ grad_p1, float* __restrict__ grad_p2,
                         float4* __restrict__ scratch /* D = 3: (N*P2) float4, else nullptr */) {
  extern __shared__ float diffs[];  // [rows_per_cta * K][DT]
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int rows = static_cast<int>(min(static_cast<int64_t>(rows_per_cta), total_rows - row0));
  const int items = rows * K;
  // UNB entries per thread at once: their indices and upstream gradients are requested first, then the
  // p2 rows they point at, then the arithmetic -- the chain idx -> p2[idx] is two dependent loads and
  // one entry per thread leaves the memory system idle most of the time
  constexpr int UNB = 4;
  for (int it0 = threadIdx.x; it0 < items; it0 += blockDim.x * UNB) {
    long long i2[UNB];
    float g[UNB];
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const int it = min(it0 + u * static_cast<int>(blockDim.x), items - 1);
      i2[u] = __ldg(idx + row0 * K + it);
      g[u] = __ldg(grad_dists + row0 * K + it);
    }
    int nn[UNB];
    int64_t rw[UNB];
    bool ok[UNB];
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const int it = it0 + u * static_cast<int>(blockDim.x);
      const int itc = min(it, items - 1);
      const int r = itc / K, k = itc - r * K;
      rw[u] = row0 + r;
      nn[u] = static_cast<int>(rw[u] / P1);
      const int i1 = static_cast<int>(rw[u] - static_cast<int64_t>(nn[u]) * P1);
      const int64_t L1v = __ldg(len1 + nn[u]), L2v = __ldg(len2 + nn[u]);  // both, unconditionally: no branch between loads
      ok[u] = (it < items) & (i1 < L1v) & (k < L2v)       // k < min(L2, K)
              & (i2[u] >= 0) & (i2[u] < P2);               // -1 = padding (ball query)
    }
    float av[UNB][DT], bv[UNB][DT];
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const float* a = p1 + rw[u] * DT;
      const float* b = p2 + (static_cast<int64_t>(nn[u]) * P2 + (ok[u] ? i2[u] : 0)) * DT;
      if (DT == 3) {  // two loads per 12-byte row instead of three
        ldg_row3(a, av[u][0], av[u][1 % DT], av[u][2 % DT]);
        ldg_row3(b, bv[u][0], bv[u][1 % DT], bv[u][2 % DT]);
      } else {
#pragma unroll
        for (int d = 0; d < DT; ++d) {
          av[u][d] = __ldg(a + d);
          bv[u][d] = __ldg(b + d);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const int it = it0 + u * static_cast<int>(blockDim.x);
      if (it >= items) continue;
      float df[DT];
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        // same float ops, same order as knn_cpu.cpp:113-122; intrinsics forbid contraction
        if (NORM == 1) df[d] = __fmul_rn(g[u], (av[u][d] > bv[u][d]) ? 1.0f : -1.0f);
        else df[d] = __fmul_rn(__fmul_rn(2.0f, g[u]), __fsub_rn(av[u][d], bv[u][d]));
        if (!ok[u]) df[d] = 0.0f;
      }
      if (ok[u]) {
        const int64_t prow = static_cast<int64_t>(nn[u]) * P2 + i2[u];
        if (DT == 3 && scratch != nullptr) {
          red_add_v4(scratch + prow, -df[0], -df[1 % DT], -df[2 % DT], 0.0f);
        } else if (DT == 4 && (reinterpret_cast<uintptr_t>(grad_p2) & 15) == 0) {
          red_add_v4(reinterpret_cast<float4*>(grad_p2) + prow, -df[0], -df[1 % DT], -df[2 % DT], -df[3 % DT]);
        } else if (DT == 2 && (reinterpret_cast<uintptr_t>(grad_p2) & 7) == 0) {
          red_add_v2(reinterpret_cast<float2*>(grad_p2) + prow, -df[0], -df[1 % DT]);
        } else {
          float* gp = grad_p2 + prow * DT;
#pragma unroll
          for (int d = 0; d < DT; ++d) atomicAdd(gp + d, -df[d]);
        }
      }
#pragma unroll
      for (int d = 0; d < DT; ++d) diffs[it * DT + d] = df[d];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < rows * DT; t += blockDim.x) {
    const int r = t / DT, d = t - r * DT;
    const float* src = diffs + static_cast<size_t>(r) * K * DT + d;
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, src[k * DT]);  // +0 for skipped entries: exact
    grad_p1[row0 * DT + t] = acc;
  }
}

// grad_p2 (rows of 3 floats) <- scratch (rows of 4 floats), flat over floats: coalesced both ways
__global__ void knn_backward_compact_kernel(const float* __restrict__ scratch, int64_t total_floats,
                                            float* __restrict__ grad_p2) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total_floats;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = e / 3;
    grad_p2[e] = scratch[row * 4 + (e - row * 3)];
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kTiledThreads = 128;
constexpr size_t kMaxSmem = 227 * 1024;

constexpr int kGenericThreads = 128;

inline void generic_layout(int D, int K, int* TP, int* lists_in_smem, size_t* smem) {
  const size_t qbytes = size_t(D) * kGenericThreads * 4;
  int tp = static_cast<int>(std::max<size_t>(8, std::min<size_t>(256, (32 * 1024) / (size_t(D) * 4))));
  size_t base = align_up(qbytes + size_t(tp) * D * 4, 8);
  while (base > 96 * 1024 && tp > 1) {
    tp /= 2;
    base = align_up(qbytes + size_t(tp) * D * 4, 8);
  }
  const size_t lbytes = size_t(K) * kGenericThreads * 8;
  *TP = tp;
  *lists_in_smem = (base + lbytes <= 160 * 1024) ? 1 : 0;
  *smem = base + (*lists_in_smem ? lbytes : 0);
}

// The exact thread-per-query kernel: the search itself for shapes no filtered path covers, the
// recomputation pass behind the tensor-core path (flags), and -- gated per cloud by dirty_bits -- the
// answer for clouds with non-finite or huge coordinates that the filtered searches skip.
int run_generic(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1, int P2,
                int D, int K, int norm, int64_t* idx, float* dists, uint64_t* glists, const unsigned char* flags,
                const unsigned* flag_count, unsigned flag_limit, const unsigned* dirty_bits, cudaStream_t st) {
  KnnGenericParams prm;
  prm.flags = flags; prm.flag_count = flag_count; prm.flag_limit = flag_limit; prm.dirty_bits = dirty_bits;
  prm.p1 = p1; prm.p2 = p2; prm.len1 = len1; prm.len2 = len2; prm.idx = idx; prm.dists = dists;
  prm.glists = glists;
  prm.P1 = P1; prm.P2 = P2; prm.D = D; prm.K = K;
  size_t smem = 0;
  generic_layout(D, K, &prm.TP, &prm.lists_in_smem, &smem);
  if (smem > 227 * 1024) return fail(POPS_ERR_UNSUPPORTED, "knn: D too large for the generic kernel");
  if (!prm.lists_in_smem && glists == nullptr)
    return fail(POPS_ERR_UNSUPPORTED, "knn: K too large for the exact fallback of this path");
  dim3 grid(static_cast<unsigned>(ceil_div(P1, kGenericThreads)), N);
  profile_begin("knn_generic", st);
  if (norm == 2) {
    auto kern = knn_generic_kernel<2, kGenericThreads>;
    POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kGenericThreads, smem, st>>>(prm);
  } else {
    auto kern = knn_generic_kernel<1, kGenericThreads>;
    POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kGenericThreads, smem, st>>>(prm);
  }
  profile_end("knn_generic", st);
  POPS_LAUNCH_OK("knn_generic_kernel");
  return POPS_OK;
}

inline int pad_points(int64_t P2) {
  return static_cast<int>((P2 + kPadPoints - 1) / kPadPoints * kPadPoints);
}

inline bool tiled_k_ok(int K) { return K <= 128; }

template <int DT, int NORM, bool EXP, int Q, int KT, int RS, int THREADS = kTiledThreads>
int launch_scan(const KnnScanParams& prm, int N, cudaStream_t st) {
  using SM = KnnSmem<DT, EXP, (KT > 0), Q, THREADS, RS>;
  constexpr int QPB = Q * THREADS;
  const size_t smem = SM::total(prm.K);
  if (smem > kMaxSmem) return fail(POPS_ERR_UNSUPPORTED, "knn: K too large for the tiled kernel");
  auto kern = knn_scan_kernel<DT, NORM, EXP, Q, THREADS, KT, RS>;
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>(ceil_div(prm.P1, QPB)), N);
  profile_begin("knn_scan", st);
  kern<<<grid, THREADS, smem, st>>>(prm);
  profile_end("knn_scan", st);
  POPS_LAUNCH_OK("knn_scan_kernel");
  return POPS_OK;
}

// D = 3, L2, K <= 32: Hilbert-ordered clouds + box-pruned search (knn_order.cu, knn_prune.cu).
// Also for small clouds: its pre-pass is a single launch up to 8192 points and its search spreads a
// cloud over one warp per 32 queries, where the tiled kernel puts 512 queries of a cloud on ONE CTA
// (measured on the reference harness's sizes, 1 cloud, K = 16, kernel + launches back to back:
// P = 100: 65 vs 73 us, P = 500: 83 vs 189 us, P = 1000: 97 vs 234 us).
inline bool use_ordered(int64_t P2, int K) {
  const int force = get_option("knn_order", -1);  // test aid
  if (K > 32) return false;
  if (force >= 0) return force != 0;
  return P2 >= 64;
}

int launch_ordered(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                   int P1, int P2, int K, int64_t* idx, float* dists, void* ws, cudaStream_t st) {
  KnnOrderBuffers ob;
  knn_order_carve(ws, N, P1, P2, &ob);
  const bool self_knn = (p1 == p2) && (len1 == len2) && (P1 == P2);
  int rc = knn_order_prepass(p1, p2, len1, len2, N, P1, P2, self_knn, ob, st);
  if (rc != POPS_OK) return rc;
  rc = knn_prune_search(ob, len1, len2, N, P1, P2, K, idx, dists, st);
  if (rc != POPS_OK) return rc;
  return run_generic(p1, p2, len1, len2, N, P1, P2, 3, K, 2, idx, dists, nullptr, nullptr, nullptr, 0, ob.maxabs_bits, st);
}

template <int DT, int NORM, bool EXP>
int launch_tiled(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                 int P1, int P2, int K, int64_t* idx, float* dists, void* ws, cudaStream_t st) {
  if (EXP && use_ordered(P2, K)) return launch_ordered(p1, p2, len1, len2, N, P1, P2, K, idx, dists, ws, st);
  const int P2pad = pad_points(P2);
  unsigned* maxabs = reinterpret_cast<unsigned*>(ws);
  float* soa = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up(size_t(N) * 4, 256));
  POPS_CUDA_OK(cudaMemsetAsync(maxabs, 0, size_t(N) * 4, st));
  {
    dim3 grid(static_cast<unsigned>(ceil_div(P2pad, 256)), N);
    knn_pack_kernel<DT, EXP><<<grid, 256, 0, st>>>(p2, len2, P2, P2pad, soa, maxabs);
    POPS_LAUNCH_OK("knn_pack_kernel");
    // p1 too: its magnitude enters the filter's error bound (EXP) and, for every variant, decides
    // whether the cloud takes the exact generic kernel instead (non-finite / huge coordinates)
    dim3 g2(static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(64, ceil_div(int64_t(P1) * DT, 1024)))), N);
    maxabs_kernel<<<g2, 256, 0, st>>>(p1, len1, P1, DT, maxabs);
    POPS_LAUNCH_OK("maxabs_kernel");
  }
  KnnScanParams prm;
  prm.p1 = p1; prm.soa = soa; prm.len1 = len1; prm.len2 = len2; prm.maxabs_bits = maxabs;
  prm.idx = idx; prm.dists = dists; prm.P1 = P1; prm.P2 = P2; prm.P2pad = P2pad; prm.K = K;
  int rc;
  if (EXP) {
    // register-merge buckets (KT = next power of two >= K)
    if (K == 1) rc = launch_scan<DT, NORM, EXP, 4, EXP ? 1 : 0, 2048>(prm, N, st);
    else if (K <= 4) rc = launch_scan<DT, NORM, EXP, 4, EXP ? 4 : 0, 2048>(prm, N, st);
    else if (K <= 16) rc = launch_scan<DT, NORM, EXP, 4, EXP ? 16 : 0, 2048>(prm, N, st);
    else if (K <= 32) rc = launch_scan<DT, NORM, EXP, EXP ? 4 : 1, EXP ? 32 : 0, EXP ? 2048 : 1024>(prm, N, st);
    else rc = launch_scan<DT, NORM, EXP, 1, 0, 1024>(prm, N, st);
  } else if (K <= 12) {
    rc = launch_scan<DT, NORM, EXP, 4, 0, 2048>(prm, N, st);
  } else {
    rc = launch_scan<DT, NORM, EXP, 1, 0, 1024>(prm, N, st);
  }
  if (rc != POPS_OK) return rc;
  return run_generic(p1, p2, len1, len2, N, P1, P2, DT, K, NORM, idx, dists, nullptr, nullptr, nullptr, 0, maxabs, st);
}

}  // namespace
}  // namespace pops

using namespace pops;

extern "C" size_t pops_knn_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K,
                                           int norm) {
  if (N <= 0 || P1 <= 0 || K <= 0) return 256;
  size_t tiled = align_up(size_t(N) * 4, 256) + size_t(N) * (D + 1) * pad_points(P2) * 4;
  if (D == 3 && norm == 2) tiled = std::max(tiled, knn_order_workspace_bytes(N, P1, P2));
  size_t generic = size_t(N) * ceil_div(P1, kGenericThreads) * K * kGenericThreads * 8;
  // the tensor-core path keeps its own buffers AND may hand flagged queries to the generic kernel
  if (knn_tc_supported(P1, P2, D, K, norm)) generic = align_up(generic, 256) + knn_tc_workspace_bytes(N, P1, P2);
  return align_up(std::max(tiled, generic), 256) + 256;
}

extern "C" int pops_knn_check_version(int version, int64_t D, int64_t K) {
  // the predicates of the reference's kernel variants (knn.cu:292-303)
  if (version == 0) return 1;
  if (version == 1) return D <= 32;
  if (version == 2) return D <= 8 && K <= 32;
  if (version == 3) return D <= 8 && K <= 4;
  return 0;
}

extern "C" int pops_knn_points_idx(const float* p1, const float* p2, const int64_t* lengths1,
                                   const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                                   int64_t D, int64_t K, int norm, int version, int64_t* idx,
                                   float* dists, void* workspace, size_t workspace_bytes,
                                   pops_stream_t stream) {
  (void)version;
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  if (N == 0 || P1 == 0 || K == 0) return POPS_OK;  // empty outputs (knn.cu:346-349)
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2 && idx && dists, "null pointer argument");
  POPS_CHECK_ARG(P2 < (int64_t(1) << 31) && P1 < (int64_t(1) << 31) && N < 65536, "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (workspace_bytes < pops_knn_workspace_bytes(N, P1, P2, D, K, norm) || !workspace)
    return fail(POPS_ERR_WORKSPACE, "knn: workspace missing or too small");
  if (P2 == 0 || D == 0) {
    POPS_CUDA_OK(cudaMemsetAsync(idx, 0, size_t(N) * P1 * K * 8, st));
    POPS_CUDA_OK(cudaMemsetAsync(dists, 0, size_t(N) * P1 * K * 4, st));
    return POPS_OK;
  }
  const int n = int(N), p1n = int(P1), p2n = int(P2), k = int(K);
#define POPS_TILED(DT, NORM, EXP)                                                             \
  if (D == DT && norm == NORM && tiled_k_ok(k)) \
    return launch_tiled<DT, NORM, EXP>(p1, p2, lengths1, lengths2, n, p1n, p2n, k, idx, dists, workspace, st);
  POPS_TILED(3, 2, true)
  POPS_TILED(3, 1, false)
  POPS_TILED(2, 2, false)
  POPS_TILED(2, 1, false)
  POPS_TILED(4, 2, false)
  POPS_TILED(4, 1, false)
  POPS_TILED(1, 2, false)
  POPS_TILED(1, 1, false)
#undef POPS_TILED
  // generic (also the exact recomputation pass behind the tensor-core path)
  const unsigned char* flags = nullptr;
  const unsigned* flag_count = nullptr;
  unsigned flag_limit = 0;
  if (knn_tc_supported(P1, P2, D, K, norm)) {
    const size_t goff = align_up(size_t(N) * ceil_div(P1, kGenericThreads) * K * kGenericThreads * 8, 256);
    unsigned char* f = nullptr;
    const int rc = knn_tc_search(p1, p2, lengths1, lengths2, n, p1n, p2n, int(D), k, idx, dists,
                                 reinterpret_cast<char*>(workspace) + goff, &f, &flag_count, &flag_limit, st);
    if (rc != POPS_OK) return rc;
    flags = f;
  }
  return run_generic(p1, p2, lengths1, lengths2, n, p1n, p2n, int(D), k, norm, idx, dists,
                     reinterpret_cast<uint64_t*>(workspace), flags, flag_count, flag_limit, nullptr, st);
}

// ---- two-phase form for host pipelines (host.py: HostKnn) ---------------------------------------
// The batch is independent per cloud, but the ordered search is preceded by a pre-pass whose ~10
// small launches cost the same for 8 clouds as for 32.  A pipeline that wants results slice by slice
// (so that their device-to-host copies overlap the search of the next slice) runs the pre-pass once
// for the whole batch and then searches ranges of clouds.
namespace {
inline bool ordered_path(int64_t P2, int64_t D, int64_t K, int norm) {
  return D == 3 && norm == 2 && tiled_k_ok(int(K)) && use_ordered(P2, int(K));
}
}  // namespace

extern "C" int pops_knn_points_prepare(const float* p1, const float* p2, const int64_t* lengths1,
                                       const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                                       int64_t D, int64_t K, int norm, void* workspace,
                                       size_t workspace_bytes, pops_stream_t stream) {
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  if (N == 0 || P1 == 0 || K == 0 || P2 == 0 || D == 0) return POPS_OK;
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2, "null pointer argument");
  POPS_CHECK_ARG(P2 < (int64_t(1) << 31) && P1 < (int64_t(1) << 31) && N < 65536, "size too large");
  if (workspace_bytes < pops_knn_workspace_bytes(N, P1, P2, D, K, norm) || !workspace)
    return fail(POPS_ERR_WORKSPACE, "knn: workspace missing or too small");
  if (!ordered_path(P2, D, K, norm)) return POPS_OK;  // nothing to share between ranges
  KnnOrderBuffers ob;
  knn_order_carve(workspace, N, P1, P2, &ob);
  const bool self_knn = (p1 == p2) && (lengths1 == lengths2) && (P1 == P2);
  return knn_order_prepass(p1, p2, lengths1, lengths2, int(N), int(P1), int(P2), self_knn, ob,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int pops_knn_points_idx_range(const float* p1, const float* p2, const int64_t* lengths1,
                                         const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                                         int64_t D, int64_t K, int norm, int version, int64_t n0,
                                         int64_t n1, int64_t* idx, float* dists, void* workspace,
                                         size_t workspace_bytes, pops_stream_t stream) {
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  POPS_CHECK_ARG(0 <= n0 && n0 <= n1 && n1 <= N, "cloud range outside the batch");
  if (n0 == n1 || P1 == 0 || K == 0) return POPS_OK;
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2 && idx && dists, "null pointer argument");
  if (workspace_bytes < pops_knn_workspace_bytes(N, P1, P2, D, K, norm) || !workspace)
    return fail(POPS_ERR_WORKSPACE, "knn: workspace missing or too small");
  int64_t* idx_r = idx + size_t(n0) * P1 * K;
  float* dists_r = dists + size_t(n0) * P1 * K;
  if (P2 == 0 || D == 0 || !ordered_path(P2, D, K, norm))  // no shared pre-pass: the range is a batch of its own
    return pops_knn_points_idx(p1 + size_t(n0) * P1 * D, p2 + size_t(n0) * P2 * D, lengths1 + n0, lengths2 + n0,
                               n1 - n0, P1, P2, D, K, norm, version, idx_r, dists_r, workspace, workspace_bytes,
                               stream);
  KnnOrderBuffers ob;
  knn_order_carve(workspace, N, P1, P2, &ob);
  const size_t nbox = size_t(knn_order_num_boxes(P2));
  ob.maxabs_bits += n0;
  ob.blocks += size_t(n0) * nbox * kBlockFloats;
  ob.boxes += size_t(n0) * nbox * 2;
  ob.qsorted += size_t(n0) * P1;
  ob.qhome += size_t(n0) * P1;
  const int rc = knn_prune_search(ob, lengths1 + n0, lengths2 + n0, int(n1 - n0), int(P1), int(P2), int(K), idx_r,
                                  dists_r, static_cast<cudaStream_t>(stream));
  if (rc != POPS_OK) return rc;
  return run_generic(p1 + size_t(n0) * P1 * D, p2 + size_t(n0) * P2 * D, lengths1 + n0, lengths2 + n0, int(n1 - n0),
                     int(P1), int(P2), 3, int(K), 2, idx_r, dists_r, nullptr, nullptr, nullptr, 0, ob.maxabs_bits,
                     static_cast<cudaStream_t>(stream));
}

// ---- both directions of a two-sided search (chamfer) ------------------------------------------------
namespace {
inline bool pair_ordered(int64_t P1, int64_t P2, int64_t D, int64_t K, int norm) {
  return ordered_path(P1, D, K, norm) && ordered_path(P2, D, K, norm) && get_option("knn_pair", 1) != 0;
}
}  // namespace

extern "C" size_t pops_knn_pair_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K,
                                                int norm) {
  const size_t a = pops_knn_workspace_bytes(N, P1, P2, D, K, norm), b = pops_knn_workspace_bytes(N, P2, P1, D, K, norm);
  return align_up(a, 256) + b;  // one region per direction (the shared sort lives in the first)
}

extern "C" int pops_knn_points_idx_pair(const float* p1, const float* p2, const int64_t* lengths1,
                                        const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                                        int64_t D, int64_t K, int norm, int64_t* idx12, float* dists12,
                                        int64_t* idx21, float* dists21, void* workspace,
                                        size_t workspace_bytes, pops_stream_t stream) {
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  if (workspace_bytes < pops_knn_pair_workspace_bytes(N, P1, P2, D, K, norm) || !workspace)
    return fail(POPS_ERR_WORKSPACE, "knn: workspace missing or too small");
  const size_t wa = align_up(pops_knn_workspace_bytes(N, P1, P2, D, K, norm), 256);
  char* wsa = reinterpret_cast<char*>(workspace);
  char* wsb = wsa + wa;
  if (N == 0 || K == 0 || P1 == 0 || P2 == 0 || D == 0 || !pair_ordered(P1, P2, D, K, norm)) {
    const int rc = pops_knn_points_idx(p1, p2, lengths1, lengths2, N, P1, P2, D, K, norm, -1, idx12, dists12, wsa, wa,
                                       stream);
    if (rc != POPS_OK) return rc;
    return pops_knn_points_idx(p2, p1, lengths2, lengths1, N, P2, P1, D, K, norm, -1, idx21, dists21, wsb,
                               workspace_bytes - wa, stream);
  }
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2 && idx12 && dists12 && idx21 && dists21, "null pointer argument");
  POPS_CHECK_ARG(P2 < (int64_t(1) << 31) && P1 < (int64_t(1) << 31) && N < 65536, "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KnnOrderBuffers a, b;
  knn_order_carve(wsa, N, P1, P2, &a);
  knn_order_carve(wsb, N, P2, P1, &b);
  b.maxabs_bits = a.maxabs_bits;
  b.bbox = a.bbox;
  int rc = knn_order_prepass_pair(p1, p2, lengths1, lengths2, int(N), int(P1), int(P2), a, b, st);
  if (rc != POPS_OK) return rc;
  rc = knn_prune_search(a, lengths1, lengths2, int(N), int(P1), int(P2), int(K), idx12, dists12, st);
  if (rc != POPS_OK) return rc;
  rc = knn_prune_search(b, lengths2, lengths1, int(N), int(P2), int(P1), int(K), idx21, dists21, st);
  if (rc != POPS_OK) return rc;
  // clouds with non-finite / huge coordinates (a.maxabs_bits spans both tensors) take the exact kernel
  rc = run_generic(p1, p2, lengths1, lengths2, int(N), int(P1), int(P2), 3, int(K), 2, idx12, dists12, nullptr, nullptr,
                   nullptr, 0, a.maxabs_bits, st);
  if (rc != POPS_OK) return rc;
  return run_generic(p2, p1, lengths2, lengths1, int(N), int(P2), int(P1), 3, int(K), 2, idx21, dists21, nullptr,
                     nullptr, nullptr, 0, a.maxabs_bits, st);
}

extern "C" size_t pops_knn_backward_workspace_bytes(int64_t N, int64_t P2, int64_t D) {
  return (D == 3 && N > 0 && P2 > 0) ? size_t(N) * size_t(P2) * 16 + 256 : 256;
}

extern "C" int pops_knn_points_backward_ws(const float* p1, const float* p2, const int64_t* lengths1,
                                           const int64_t* lengths2, const int64_t* idx,
                                           const float* grad_dists, int64_t N, int64_t P1, int64_t P2,
                                           int64_t D, int64_t K, int norm, float* grad_p1,
                                           float* grad_p2, void* workspace, size_t workspace_bytes,
                                           pops_stream_t stream) {
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t total = size_t(N) * P1 * D;
  const bool rows_path = D >= 1 && D <= 4 && K >= 1 && K <= 1024 && get_option("knn_backward_rows", 1) != 0;
  float4* scratch = nullptr;
  if (rows_path && D == 3 && workspace != nullptr &&
      workspace_bytes >= pops_knn_backward_workspace_bytes(N, P2, D) &&
      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && total > 0 && P2 > 0)
    scratch = reinterpret_cast<float4*>(workspace);
  if (N * P2 * D > 0) {
    POPS_CHECK_ARG(grad_p2, "null grad_p2");
    if (scratch) POPS_CUDA_OK(cudaMemsetAsync(scratch, 0, size_t(N) * P2 * 16, st));
    else POPS_CUDA_OK(cudaMemsetAsync(grad_p2, 0, size_t(N) * P2 * D * 4, st));
  }
  if (total == 0) return POPS_OK;
  POPS_CHECK_ARG(p1 && lengths1 && lengths2 && grad_p1, "null pointer argument");
  if (K == 0 || P2 == 0) {
    POPS_CUDA_OK(cudaMemsetAsync(grad_p1, 0, total * 4, st));
    return POPS_OK;
  }
  POPS_CHECK_ARG(p2 && idx && grad_dists, "null pointer argument");
  profile_begin("knn_backward", st);
  if (rows_path) {
    const int rows_per_cta = std::max(1, 1024 / int(K));
    const int64_t rows = N * P1;
    const unsigned grid = static_cast<unsigned>(ceil_div(rows, rows_per_cta));
    const size_t smem = size_t(rows_per_cta) * K * D * 4;
#define POPS_BWD(NORM, DT)                                                                                  \
  knn_backward_rows_kernel<NORM, DT><<<grid, 256, smem, st>>>(p1, p2, lengths1, lengths2, idx, grad_dists,  \
                                                              rows, int(P1), int(P2), int(K), rows_per_cta, \
                                                              grad_p1, grad_p2, scratch)
#define POPS_BWD_D(NORM)                 \
  switch (D) {                           \
    case 1: POPS_BWD(NORM, 1); break;    \
    case 2: POPS_BWD(NORM, 2); break;    \
    case 3: POPS_BWD(NORM, 3); break;    \
    default: POPS_BWD(NORM, 4); break;   \
  }
    if (norm == 2) { POPS_BWD_D(2) } else { POPS_BWD_D(1) }
#undef POPS_BWD_D
#undef POPS_BWD
    POPS_LAUNCH_OK("knn_backward_rows_kernel");
    if (scratch) {
      const int64_t floats = N * P2 * 3;
      const int blocks = int(std::min<int64_t>(ceil_div(floats, 256), int64_t(num_sms()) * 16));
      knn_backward_compact_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(scratch), floats, grad_p2);
      POPS_LAUNCH_OK("knn_backward_compact_kernel");
    }
    profile_end("knn_backward", st);
    return POPS_OK;
  }
  const int threads = 256;
  const int blocks = int(std::min<int64_t>(ceil_div(int64_t(total), threads), int64_t(num_sms()) * 16));
  if (norm == 2)
    knn_backward_kernel<2><<<blocks, threads, 0, st>>>(p1, p2, lengths1, lengths2, idx, grad_dists,
                                                       int(N), int(P1), int(P2), int(D), int(K),
                                                       grad_p1, grad_p2);
  else
    knn_backward_kernel<1><<<blocks, threads, 0, st>>>(p1, p2, lengths1, lengths2, idx, grad_dists,
                                                       int(N), int(P1), int(P2), int(D), int(K),
                                                       grad_p1, grad_p2);
  profile_end("knn_backward", st);
  POPS_LAUNCH_OK("knn_backward_kernel");
  return POPS_OK;
}

extern "C" int pops_knn_points_backward(const float* p1, const float* p2, const int64_t* lengths1,
                                        const int64_t* lengths2, const int64_t* idx,
                                        const float* grad_dists, int64_t N, int64_t P1, int64_t P2,
                                        int64_t D, int64_t K, int norm, float* grad_p1,
                                        float* grad_p2, pops_stream_t stream) {
  return pops_knn_points_backward_ws(p1, p2, lengths1, lengths2, idx, grad_dists, N, P1, P2, D, K, norm, grad_p1,
                                     grad_p2, nullptr, 0, stream);
}
