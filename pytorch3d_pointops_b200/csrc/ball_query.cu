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
//   D == 3: tile is SoA (x[],y[],z[],w=|p|^2) and points are first screened 4 at a time with the
//           expanded-form filter  s = w - 2 q.p  against  (r2 - |q|^2) + E  (same error bound as
//           the KNN filter); only groups that pass are evaluated exactly.  The exact
//           value alone decides membership.
//   other D: exact distance on an AoS tile.
#include <cfloat>

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
      const float T = __fadd_rn(__fsub_rn(prm.radius2, qq), E);
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

}  // namespace pops

using namespace pops;

extern "C" size_t pops_ball_query_workspace_bytes(int64_t, int64_t, int64_t, int64_t, int64_t) {
  return 256;
}

extern "C" int pops_ball_query(const float* p1, const float* p2, const int64_t* lengths1,
                               const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2,
                               int64_t D, int64_t K, float radius, int64_t* idx, float* dists,
                               void* workspace, size_t workspace_bytes, pops_stream_t stream) {
  (void)workspace;
  (void)workspace_bytes;
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0 && K >= 0, "negative size");
  if (N == 0 || P1 == 0 || K == 0) return POPS_OK;
  POPS_CHECK_ARG(p1 && p2 && lengths1 && lengths2 && idx && dists, "null pointer argument");
  POPS_CHECK_ARG(P2 < (int64_t(1) << 31) && P1 < (int64_t(1) << 31) && N < 65536, "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BqParams prm;
  prm.p1 = p1; prm.p2 = p2; prm.len1 = lengths1; prm.len2 = lengths2; prm.idx = idx; prm.dists = dists;
  prm.P1 = int(P1); prm.P2 = int(P2); prm.D = int(D); prm.K = int(K);
  prm.radius2 = radius * radius;  // f32 product, ball_query_cpu.cpp:26
  dim3 grid(static_cast<unsigned>(ceil_div(P1, kBqThreads)), static_cast<unsigned>(N));
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
