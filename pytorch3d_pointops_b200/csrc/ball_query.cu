// Ball query for sm_100a.
//
// Replaces csrc/ball_query/{ball_query.cu,ball_query_cpu.cpp}.  Contract = the reference CPU path
// (ball_query_cpu.cpp:12-54): for each query the FIRST K points of p2, in index order, whose
// unfused float32 squared distance is strictly below radius2 = fl(radius*radius); idx padded
// with -1, dists with 0.
//
// Kernel: one thread per query, p2 streamed through shared memory in tiles; every thread scans
// the tile in index order and stops once it holds K hits; the CTA leaves the tile loop as soon
// as all of its queries are complete (__syncthreads_or), which is where the time goes for
// dense clouds (sequential-scan semantics make the work data dependent).
//   D == 3: ball_query_scan_kernel -- the structure of the KNN scan (knn.cu): every thread owns 4
//           queries; the tile is SoA (x[],y[],z[],w=|p|^2) built on the fly; per group of 4 points
//           and per query 6 FFMA2 evaluate the expanded form  s = w - 2 q.p  against
//           (r2 - |q|^2) + E  (same error bound as the KNN filter) and a group that may
//           hold a hit is appended, predicated, to the query's candidate buffer.  Buffers are
//           drained in warp-converged flushes, in index order: exact unfused distance, strict
//           d < r2, hits written straight to the outputs until the query holds K.  The exact value
//           alone decides membership.  (ball_query_d3_kernel, thread per query with the recheck
//           inside the loop, is kept for small inputs.)
//   other D: exact distance on an AoS tile.
#include <cfloat>

#include "bq_prune.cuh"
#include "common.cuh"

namespace pops {

constexpr int kBqThreads = 256;

struct BqParams {
  const float* p1;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  int64_t* idx;
  float* dists;
  int P1, P2, D, K, TP;
  float radius2;
  const unsigned* taken;  // scan kernel only: taken[n] != 0 -> bq_prune_kernel answers cloud n (nullptr: none)
};

// exact sequential scan, any D (AoS tile)
__global__ void __launch_bounds__(kBqThreads)
ball_query_generic_kernel(const BqParams prm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = blockIdx.y, tid = threadIdx.x, qi = blockIdx.x * kBqThreads + tid;
  const int D = prm.D, K = prm.K, TP = prm.TP;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  float* qsm = reinterpret_cast<float*>(smem);          // [D][threads]
  float* tile = qsm + static_cast<size_t>(D) * kBqThreads;  // [TP][D]
  const bool valid = qi < L1;
  for (int d = 0; d < D; ++d)
    qsm[d * kBqThreads + tid] = valid ? prm.p1[(static_cast<size_t>(n) * prm.P1 + qi) * D + d] : 0.0f;
  int64_t* oi = prm.idx + (static_cast<size_t>(n) * prm.P1 + (qi < prm.P1 ? qi : 0)) * K;
  float* od = prm.dists + (static_cast<size_t>(n) * prm.P1 + (qi < prm.P1 ? qi : 0)) * K;
  int count = 0;
  const float* p2n = prm.p2 + static_cast<size_t>(n) * prm.P2 * D;
  for (int j0 = 0; j0 < L2; j0 += TP) {
    const int pts = min(TP, L2 - j0);
    if (!__syncthreads_or(valid && count < K)) break;
    for (int e = tid; e < pts * D; e += kBqThreads) tile[e] = p2n[static_cast<size_t>(j0) * D + e];
    __syncthreads();
    if (valid) {
      for (int jl = 0; jl < pts && count < K; ++jl) {
        const float* pt = tile + jl * D;
        float d2 = 0.0f;
        for (int dd = 0; dd < D; ++dd) d2 = __fadd_rn(d2, dist_term<2>(qsm[dd * kBqThreads + tid], pt[dd]));
        if (d2 < prm.radius2) {
          oi[count] = j0 + jl;
          od[count] = d2;
          ++count;
        }
      }
    }
  }
  if (qi < prm.P1)
    for (int k = count; k < K; ++k) {
      oi[k] = -1;
      od[k] = 0.0f;
    }
}

// D == 3: SoA tile built on the fly + expanded-form group filter.
__global__ void __launch_bounds__(kBqThreads)
ball_query_d3_kernel(const BqParams prm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = blockIdx.y, tid = threadIdx.x, qi = blockIdx.x * kBqThreads + tid;
  const int K = prm.K, TP = prm.TP;  // TP multiple of 4
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  float* tx = reinterpret_cast<float*>(smem);
  float* ty = tx + TP;
  float* tz = ty + TP;
  float* tw = tz + TP;
  __shared__ float s_maxabs;
  const bool valid = qi < L1;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (valid) {
    const float* q = prm.p1 + (static_cast<size_t>(n) * prm.P1 + qi) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  const float ax = -2.0f * qx, ay = -2.0f * qy, az = -2.0f * qz;
  const float qq = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
  float qmax = fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz)));
  int64_t* oi = prm.idx + (static_cast<size_t>(n) * prm.P1 + (qi < prm.P1 ? qi : 0)) * K;
  float* od = prm.dists + (static_cast<size_t>(n) * prm.P1 + (qi < prm.P1 ? qi : 0)) * K;
  int count = 0;
  const float* p2n = prm.p2 + static_cast<size_t>(n) * prm.P2 * 3;
  const float inf = __int_as_float(0x7f800000);
  for (int j0 = 0; j0 < L2; j0 += TP) {
    const int pts = min(TP, L2 - j0);
    const int pts4 = (pts + 3) & ~3;
    if (!__syncthreads_or(valid && count < K)) break;
    if (tid == 0) s_maxabs = 0.0f;
    __syncthreads();
    float m = 0.0f;
    for (int jl = tid; jl < pts4; jl += kBqThreads) {
      float x = 0.f, y = 0.f, z = 0.f, w = inf;
      if (jl < pts) {
        const float* p = p2n + static_cast<size_t>(j0 + jl) * 3;
        x = p[0]; y = p[1]; z = p[2];
        w = fmaf(z, z, fmaf(y, y, x * x));
        m = fmaxf(m, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
      }
      tx[jl] = x; ty[jl] = y; tz[jl] = z; tw[jl] = w;
    }
    m = warp_max(m);
    if ((tid & 31) == 0) atomicMax(reinterpret_cast<int*>(&s_maxabs), __float_as_int(m));
    __syncthreads();
    if (valid && count < K) {
      const float M = fmaxf(s_maxabs, qmax);
      const float E = fmaf(M * M, 1.52587890625e-05f, 1e-37f);
      // beyond 1e18 (or +inf) the expanded form proves nothing: every group goes to the exact test.
      // NaN coordinates need no care: fmaxf / fminf drop them and the exact test d2 < r2 is false,
      // as in the reference (ball_query_cpu.cpp:44)
      const float T = (M < 1e18f) ? __fadd_rn(__fsub_rn(prm.radius2, qq), E) : inf;
      for (int g = 0; g < pts4 / 4 && count < K; ++g) {
        const float4 X = reinterpret_cast<const float4*>(tx)[g];
        const float4 Y = reinterpret_cast<const float4*>(ty)[g];
        const float4 Z = reinterpret_cast<const float4*>(tz)[g];
        const float4 W = reinterpret_cast<const float4*>(tw)[g];
        const float s0 = fmaf(az, Z.x, fmaf(ay, Y.x, fmaf(ax, X.x, W.x)));
        const float s1 = fmaf(az, Z.y, fmaf(ay, Y.y, fmaf(ax, X.y, W.y)));
        const float s2 = fmaf(az, Z.z, fmaf(ay, Y.z, fmaf(ax, X.z, W.z)));
        const float s3 = fmaf(az, Z.w, fmaf(ay, Y.w, fmaf(ax, X.w, W.w)));
        if (fminf(fminf(s0, s1), fminf(s2, s3)) <= T) {
          const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int jl = g * 4 + i;
            const float dx = __fsub_rn(qx, xs[i]), dy = __fsub_rn(qy, ys[i]), dz = __fsub_rn(qz, zs[i]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < prm.radius2 && jl < pts && count < K) {
              oi[count] = j0 + jl;
              od[count] = d2;
              ++count;
            }
          }
        }
      }
    }
  }
  if (qi < prm.P1)
    for (int k = count; k < K; ++k) {
      oi[k] = -1;
      od[k] = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// D == 3, buffered: 4 queries per thread, candidate groups flushed out of the dense loop
// ---------------------------------------------------------------------------------------------
constexpr int kBqsThreads = 128;
constexpr int kBqsQ = 4;
constexpr int kBqsTile = 1024;   // points per tile
constexpr int kBqsCap = 16;      // candidate groups a query buffers between flushes
constexpr int kBqsChunk = 4;     // groups between overflow checks

// Drain one query's candidate buffer (group ids within the tile, in scan order = index order).
// Not inlined; warp-converged call, divergent inside (rare).  Returns the new hit count.
__device__ __noinline__ int bq_flush_one(const float* tile, const unsigned short* cand_col, int c_end, float qx,
                                         float qy, float qz, float r2, int j0, int L2, int K, int count,
                                         int64_t* oi, float* od) {
  constexpr int QPB = kBqsQ * kBqsThreads;
  for (int c = 0; c < c_end && count < K; ++c) {
    const int g = cand_col[c * QPB];
    const float4 X = reinterpret_cast<const float4*>(tile)[g];
    const float4 Y = reinterpret_cast<const float4*>(tile + kBqsTile)[g];
    const float4 Z = reinterpret_cast<const float4*>(tile + 2 * kBqsTile)[g];
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + g * 4 + i;
      const float dx = __fsub_rn(qx, xs[i]), dy = __fsub_rn(qy, ys[i]), dz = __fsub_rn(qz, zs[i]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 < r2 && j < L2 && count < K) {
        oi[count] = j;
        od[count] = d2;
        ++count;
      }
    }
  }
  return count;
}

__global__ void __launch_bounds__(kBqsThreads, 3)
ball_query_scan_kernel(const BqParams prm) {
  constexpr int Q = kBqsQ, THREADS = kBqsThreads, QPB = Q * THREADS, RS = kBqsTile;
  extern __shared__ __align__(16) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);                                     // x,y,z,w rows of RS (+ pad)
  unsigned short* cand = reinterpret_cast<unsigned short*>(smem + (4 * RS + 16) * 4);  // [kBqsCap][QPB]
  __shared__ float s_maxabs;
  const int n = blockIdx.y, tid = threadIdx.x, q_base = blockIdx.x * QPB;
  const int K = prm.K;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  const float INF = __int_as_float(0x7f800000);
  const float r2 = prm.radius2;
  if (prm.taken != nullptr && prm.taken[n] != 0u) return;  // answered by bq_prune_kernel

  float q[Q][3], a[Q][3], qq[Q], qmax[Q], T[Q];
  int count[Q];
  uint32_t cw[Q];
  bool valid[Q];
  constexpr uint32_t CSTRIDE = QPB * 2;
  const uint32_t cand_base = smem_u32(cand) + static_cast<uint32_t>(tid) * 2u;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const int qi = q_base + tid + t * THREADS;
    valid[t] = qi < L1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      q[t][d] = valid[t] ? prm.p1[(static_cast<size_t>(n) * prm.P1 + qi) * 3 + d] : 0.0f;
      a[t][d] = -2.0f * q[t][d];
    }
    qq[t] = fmaf(q[t][2], q[t][2], fmaf(q[t][1], q[t][1], q[t][0] * q[t][0]));
    qmax[t] = fmaxf(fabsf(q[t][0]), fmaxf(fabsf(q[t][1]), fabsf(q[t][2])));
    count[t] = 0;
    T[t] = -INF;
    cw[t] = cand_base + static_cast<uint32_t>(t) * (THREADS * 2u);
  }
  const float* p2n = prm.p2 + static_cast<size_t>(n) * prm.P2 * 3;
  const uint32_t cw_limit = cand_base + static_cast<uint32_t>(kBqsCap - kBqsChunk) * CSTRIDE;

  auto flush = [&](int t, int j0) {
    const uint32_t base = cand_base + static_cast<uint32_t>(t) * (THREADS * 2u);
    const int c_end = static_cast<int>((cw[t] - base) / CSTRIDE);
    cw[t] = base;
    if (!__any_sync(0xffffffffu, c_end > 0)) return;
    const int qi = q_base + tid + t * THREADS;
    const size_t row = static_cast<size_t>(n) * prm.P1 + (qi < prm.P1 ? qi : 0);
    count[t] = bq_flush_one(tile, cand + tid + t * THREADS, c_end, q[t][0], q[t][1], q[t][2], r2, j0, L2, K,
                            count[t], prm.idx + row * K, prm.dists + row * K);
    if (count[t] >= K) T[t] = -INF;  // complete: never buffers again
  };

  for (int j0 = 0; j0 < L2; j0 += RS) {
    const int pts = min(RS, L2 - j0);
    const int ngroups = (pts + 4 * kBqsChunk - 1) / (4 * kBqsChunk) * kBqsChunk;  // whole chunks; padded with sentinels
    bool active = false;
#pragma unroll
    for (int t = 0; t < Q; ++t) active = active || (valid[t] && count[t] < K);
    if (!__syncthreads_or(active ? 1 : 0)) break;  // every query of the CTA holds K hits
    if (tid == 0) s_maxabs = 0.0f;
    __syncthreads();
    float m = 0.0f;
    for (int jl = tid; jl < ngroups * 4; jl += THREADS) {  // (the scan's last prefetch reads one group past: in bounds, unused)
      float x = 0.f, y = 0.f, z = 0.f, w = INF;
      if (jl < pts) {
        const float* p = p2n + static_cast<size_t>(j0 + jl) * 3;
        x = p[0]; y = p[1]; z = p[2];
        w = fmaf(z, z, fmaf(y, y, x * x));
        m = fmaxf(m, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
      }
      tile[jl] = x; tile[RS + jl] = y; tile[2 * RS + jl] = z; tile[3 * RS + jl] = w;
    }
    m = warp_max(m);
    if ((tid & 31) == 0) atomicMax(reinterpret_cast<int*>(&s_maxabs), __float_as_int(m));
    __syncthreads();
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const float M = fmaxf(s_maxabs, qmax[t]);
      const float E = fmaf(M * M, 1.52587890625e-05f, 1e-37f);
      // beyond 1e18 (or +inf) the expanded form proves nothing: every group goes to the exact test
      // (NaN: dropped by fmaxf / fminf, and d2 < r2 is false for it, as in ball_query_cpu.cpp:44)
      T[t] = (valid[t] && count[t] < K) ? ((M < 1e18f) ? __fadd_rn(__fsub_rn(r2, qq[t]), E) : INF) : -INF;
    }
    const float4* tp = reinterpret_cast<const float4*>(tile);
    float4 Xc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) Xc[r] = tp[r * (RS / 4)];
    for (int g = 0; g < ngroups; g += kBqsChunk) {
#pragma unroll
      for (int c = 0; c < kBqsChunk; ++c) {
        float4 Xn[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) Xn[r] = tp[r * (RS / 4) + g + c + 1];
        const unsigned short g16 = static_cast<unsigned short>(g + c);
#pragma unroll
        for (int t = 0; t < Q; ++t) {
          float2 s01 = make_float2(Xc[3].x, Xc[3].y), s23 = make_float2(Xc[3].z, Xc[3].w);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float2 ad = make_float2(a[t][d], a[t][d]);
            s01 = __ffma2_rn(ad, make_float2(Xc[d].x, Xc[d].y), s01);
            s23 = __ffma2_rn(ad, make_float2(Xc[d].z, Xc[d].w), s23);
          }
          const float mn = fminf(fminf(s01.x, s01.y), fminf(s23.x, s23.y));
          if (mn <= T[t]) {  // predicated: one STS.U16 + one IADD
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(cw[t]), "h"(g16) : "memory");
            cw[t] += CSTRIDE;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) Xc[r] = Xn[r];
      }
      uint32_t mx = cw[0];
#pragma unroll
      for (int t = 1; t < Q; ++t) mx = max(mx, cw[t] - static_cast<uint32_t>(t) * (THREADS * 2u));
      if (g + kBqsChunk >= ngroups || __any_sync(0xffffffffu, mx > cw_limit)) {
#pragma unroll
        for (int t = 0; t < Q; ++t) flush(t, j0);
      }
    }
    __syncthreads();  // everyone is done with the tile
  }
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const int qi = q_base + tid + t * THREADS;
    if (qi >= prm.P1) continue;
    int64_t* oi = prm.idx + (static_cast<size_t>(n) * prm.P1 + qi) * K;
    float* od = prm.dists + (static_cast<size_t>(n) * prm.P1 + qi) * K;
    for (int k = count[t]; k < K; ++k) {
      oi[k] = -1;
      od[k] = 0.0f;
    }
  }
}

}  // namespace pops

using namespace pops;

namespace {
// shapes on which a cloud may take the Hilbert-ordered search (bq_prune.cu); the choice per cloud is made
// on the device (bq_cloud_spatial)
bool bq_spatial_shape(int64_t P1, int64_t P2, int64_t D, int64_t K) {
  return D == 3 && P1 >= 1024 && P2 >= kBqSpatialMinPoints && K <= kBqSpatialMaxK && get_option("bq_spatial", -1) != 0 &&
         get_option("bq_scan", 1) != 0;
}
}  // namespace

extern "C" size_t pops_ball_query_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K) {
  if (N > 0 && bq_spatial_shape(P1, P2, D, K)) return knn_order_workspace_bytes(N, P1, P2);
  return 256;
}

extern "C" int pops_ball_query(const float* p1, const float* p2, const int64_t* lengths1,
                               const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                               int64_t D, int64_t K, float radius, int64_t* idx, float* dists,
                               void* workspace, size_t workspace_bytes, pops_stream_t stream) {
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  if (N == 0 || P1 == 0 || K == 0) return POPS_OK;
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2 && idx && dists, "null pointer argument");
  POPS_CHECK_ARG(P2 < (int64_t(1) << 31) && P1 < (int64_t(1) << 31) && N < 65536, "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BqParams prm;
  prm.p1 = p1; prm.p2 = p2; prm.len1 = lengths1; prm.len2 = lengths2; prm.idx = idx; prm.dists = dists;
  prm.P1 = int(P1); prm.P2 = int(P2); prm.D = int(D); prm.K = int(K);
  prm.radius2 = radius * radius;  // f32 product, ball_query_cpu.cpp:26
  prm.taken = nullptr;
  if (bq_spatial_shape(P1, P2, D, K)) {
    POPS_CHECK_ARG(workspace && workspace_bytes >= knn_order_workspace_bytes(N, P1, P2), "ball_query: workspace too small");
    KnnOrderBuffers ob;
    knn_order_carve(workspace, N, P1, P2, &ob);
    const bool self = (p1 == p2) && (lengths1 == lengths2) && (P1 == P2);
    int rc = knn_order_prepass(p1, p2, lengths1, lengths2, int(N), int(P1), int(P2), self, ob, st);
    if (rc != POPS_OK) return rc;
    unsigned* flags = ob.keys_in;  // N words of the pre-pass scratch, free again once the order is built
    profile_begin("ball_query", st);
    rc = bq_prune_search(ob, p2, lengths1, lengths2, int(N), int(P1), int(P2), int(K), radius, prm.radius2,
                         get_option("bq_spatial", -1), flags, idx, dists, st);
    if (rc != POPS_OK) return rc;
    prm.taken = flags;
    prm.TP = kBqsTile;
    const size_t smem = size_t(4 * kBqsTile + 16) * 4 + size_t(kBqsCap) * kBqsQ * kBqsThreads * 2;
    POPS_CUDA_OK(cudaFuncSetAttribute(ball_query_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    dim3 sgrid(static_cast<unsigned>(ceil_div(P1, kBqsQ * kBqsThreads)), static_cast<unsigned>(N));
    ball_query_scan_kernel<<<sgrid, kBqsThreads, smem, st>>>(prm);
    profile_end("ball_query", st);
    POPS_LAUNCH_OK("ball_query_scan_kernel");
    return POPS_OK;
  }
  dim3 grid(static_cast<unsigned>(ceil_div(P1, kBqThreads)), static_cast<unsigned>(N));
  if (D == 3 && P1 >= 1024 && get_option("bq_scan", 1)) {
    prm.TP = kBqsTile;
    const size_t smem = size_t(4 * kBqsTile + 16) * 4 + size_t(kBqsCap) * kBqsQ * kBqsThreads * 2;
    POPS_CUDA_OK(cudaFuncSetAttribute(ball_query_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    dim3 sgrid(static_cast<unsigned>(ceil_div(P1, kBqsQ * kBqsThreads)), static_cast<unsigned>(N));
    profile_begin("ball_query", st);
    ball_query_scan_kernel<<<sgrid, kBqsThreads, smem, st>>>(prm);
    profile_end("ball_query", st);
    POPS_LAUNCH_OK("ball_query_scan_kernel");
    return POPS_OK;
  }
  if (D == 3) {
    prm.TP = 2048;
    const size_t smem = size_t(4) * prm.TP * 4;
    profile_begin("ball_query", st);
    ball_query_d3_kernel<<<grid, kBqThreads, smem, st>>>(prm);
    profile_end("ball_query", st);
    POPS_LAUNCH_OK("ball_query_d3_kernel");
    return POPS_OK;
  }
  const size_t qbytes = size_t(D) * kBqThreads * 4;
  int tp = int(std::max<size_t>(4, std::min<size_t>(512, (32 * 1024) / std::max<size_t>(1, size_t(D) * 4))));
  const size_t smem = qbytes + size_t(tp) * D * 4;
  if (smem > 200 * 1024) return fail(POPS_ERR_UNSUPPORTED, "ball_query: D too large");
  prm.TP = tp;
  POPS_CUDA_OK(cudaFuncSetAttribute(ball_query_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  ball_query_generic_kernel<<<grid, kBqThreads, smem, st>>>(prm);
  POPS_LAUNCH_OK("ball_query_generic_kernel");
  return POPS_OK;
}
