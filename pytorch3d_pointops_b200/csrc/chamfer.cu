// Fused chamfer post-processing for sm_100a.
//
// After the K=1 nearest-neighbour search, the reference's chamfer_distance runs ~20 small torch
// kernels per direction (mask, weights, knn_gather, cosine_similarity = norm/clamp/div/mul/sum,
// abs, 1-x, masked fill, sum, divide; functions/chamfer.py:114-189) and their autograd mirror
// images.  Here one forward and one backward kernel per direction do all of it:
//
//   forward   per point: cham = dist * w_n; per feature f: 1 - |cos(xf[i], yf[idx[i]])| * w_n, masked
//             for i >= lengths1[n]; reduced over the cloud's points (sum | mean | max | none).
//   backward  per point: grad_x += 2 g (x - y[idx]) (or g*sign for L1), grad_y[idx] -= same
//             (knn_cpu.cpp:113-122), and the cosine chain rule into grad_xf / grad_yf[idx].
//
// cosine_similarity follows ATen's formulation (dim=2, eps=1e-6):
//   cos = sum_k (a_k / max(|a|, eps)) * (b_k / max(|b|, eps)).
// HBM bound.  Forward: one CLUSTER of kChamferCluster CTAs per cloud; every CTA reduces its slice of
// the points, pushes its partials to rank 0 through distributed shared memory, and rank 0 adds them
// in rank order (deterministic).  Backward: no reduction at all -> flat grid over the points.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace pops {

constexpr int kChamferMaxFeats = 8;
constexpr int kChamferThreads = 512;
constexpr int kChamferCluster = 8;  // CTAs per cloud in the forward kernel (portable cluster size)
constexpr int kBwdThreads = 256;

__device__ __forceinline__ uint32_t ch_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void ch_cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store one 32-bit value into the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ void ch_st_remote(void* local_smem_ptr, uint32_t rank, uint32_t value) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(value) : "memory");
}

struct ChamferFeat {
  const float* xf[kChamferMaxFeats];
  const float* yf[kChamferMaxFeats];
  float* gxf[kChamferMaxFeats];
  float* gyf[kChamferMaxFeats];
  int chans[kChamferMaxFeats];
  int num;
};

enum { kRedNone = 0, kRedSum = 1, kRedMean = 2, kRedMax = 3 };

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  v = (lane < nw) ? sm[lane] : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cos(a, b) with ATen's eps clamping; also returns the clamped norms
__device__ __forceinline__ float cosine(const float* a, const float* b, int C, float* na_out, float* nb_out,
                                        bool b_zero) {
  float sa = 0.f, sb = 0.f;
  for (int k = 0; k < C; ++k) {
    const float av = a[k], bv = b_zero ? 0.0f : b[k];
    sa = fmaf(av, av, sa);
    sb = fmaf(bv, bv, sb);
  }
  const float na = fmaxf(sqrtf(sa), 1e-6f), nb = fmaxf(sqrtf(sb), 1e-6f);
  float c = 0.f;
  for (int k = 0; k < C; ++k) {
    const float bv = b_zero ? 0.0f : b[k];
    c = fmaf(a[k] / na, bv / nb, c);
  }
  *na_out = na;
  *nb_out = nb;
  return c;
}

__global__ void __launch_bounds__(kChamferThreads)
chamfer_fwd_kernel(const float* __restrict__ dists, const int64_t* __restrict__ idx,
                   const int64_t* __restrict__ len1, const int64_t* __restrict__ len2,
                   const float* __restrict__ weights, int P1, int P2, ChamferFeat ft, int reduction,
                   int abs_cosine, int N, float* __restrict__ cham_out, float* __restrict__ feat_out,
                   int64_t* __restrict__ argmax_out) {
  constexpr int C = kChamferCluster;
  __shared__ float sm[32];
  __shared__ int smi[32];
  // rank 0 only: partials of every CTA of the cluster (slot 0: chamfer term / max value, 1..8: features)
  __shared__ float part[C][1 + kChamferMaxFeats];
  __shared__ int parti[C];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int rank = static_cast<int>(ch_cluster_rank());
  ch_cluster_barrier();  // every CTA of the cluster is running before any of them writes a peer's shared memory
  int64_t L1l = len1[n], L2l = len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
  const bool y_empty = L2l <= 0;
  const float w = weights ? weights[n] : 1.0f;
  const float* dn = dists + static_cast<size_t>(n) * P1;
  const int64_t* in = idx + static_cast<size_t>(n) * P1;
  // this CTA's contiguous slice of the cloud's points
  const int per = (P1 + C - 1) / C;
  const int i0 = rank * per, i1 = min(P1, i0 + per);

  // ---- chamfer term -----------------------------------------------------------------------------
  if (reduction == kRedNone) {
    for (int i = i0 + tid; i < i1; i += kChamferThreads)
      cham_out[static_cast<size_t>(n) * P1 + i] = (i < L1) ? dn[i] * w : 0.0f;
  } else if (reduction == kRedMax) {
    float best = -FLT_MAX;
    int bi = 0x7fffffff;
    for (int i = i0 + tid; i < i1; i += kChamferThreads) {
      const float v = (i < L1) ? dn[i] * w : 0.0f;  // padded points count as 0, as in the reference
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { sm[tid >> 5] = best; smi[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      best = (tid < (kChamferThreads >> 5)) ? sm[tid] : -FLT_MAX;
      bi = (tid < (kChamferThreads >> 5)) ? smi[tid] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) {
        ch_st_remote(&part[rank][0], 0, __float_as_uint(best));
        ch_st_remote(&parti[rank], 0, static_cast<uint32_t>(bi));
      }
    }
    __syncthreads();
  } else {
    float s = 0.f;
    for (int i = i0 + tid; i < min(i1, L1); i += kChamferThreads) s += dn[i] * w;
    s = block_sum(s, sm);
    if (tid == 0) ch_st_remote(&part[rank][0], 0, __float_as_uint(s));
  }

  // ---- feature terms ----------------------------------------------------------------------------
  for (int f = 0; f < ft.num; ++f) {
    const int Cf = ft.chans[f];
    const float* xf = ft.xf[f] + static_cast<size_t>(n) * P1 * Cf;
    const float* yf = ft.yf[f] + static_cast<size_t>(n) * P2 * Cf;
    float s = 0.f;
    for (int i = i0 + tid; i < i1; i += kChamferThreads) {
      float fd = 0.0f;
      if (i < L1) {
        float na, nb;
        int64_t j = in[i];
        const bool no_nb = y_empty || j < 0 || j >= P2;  // an index the search never returns: treat as "no neighbour", never dereference
        if (no_nb) j = 0;
        const float c = cosine(xf + static_cast<size_t>(i) * Cf, yf + static_cast<size_t>(j) * Cf, Cf, &na, &nb, no_nb);
        fd = (1.0f - (abs_cosine ? fabsf(c) : c)) * w;
      }
      if (reduction == kRedNone) feat_out[(static_cast<size_t>(f) * N + n) * P1 + i] = fd;
      s += fd;
    }
    if (reduction != kRedNone) {
      s = block_sum(s, sm);
      if (tid == 0) ch_st_remote(&part[rank][1 + f], 0, __float_as_uint(s));
    }
  }

  // ---- rank 0 combines the CTAs' partials in rank order ---------------------------------------------
  ch_cluster_barrier();
  if (rank == 0 && tid == 0 && reduction != kRedNone) {
    if (reduction == kRedMax) {
      float best = part[0][0];
      int bi = parti[0];
      for (int r = 1; r < C; ++r)
        if (part[r][0] > best || (part[r][0] == best && parti[r] < bi)) { best = part[r][0]; bi = parti[r]; }
      cham_out[n] = best;
      argmax_out[n] = bi;
    } else {
      float s = 0.f;
      for (int r = 0; r < C; ++r) s += part[r][0];
      cham_out[n] = (reduction == kRedMean) ? s / static_cast<float>(L1 > 0 ? L1 : 1) : s;
    }
    for (int f = 0; f < ft.num; ++f) {
      float s = 0.f;
      for (int r = 0; r < C; ++r) s += part[r][1 + f];
      feat_out[static_cast<size_t>(f) * N + n] = (reduction == kRedMean) ? s / static_cast<float>(L1 > 0 ? L1 : 1) : s;
    }
  }
}

// grad buffers must be zero-filled by the caller (the host function does it unless `acc`: then they
// hold the other direction's gradient and this launch adds to it -- the per-point stores become
// read-modify-writes, which no other thread of this launch touches: its atomics go to the other
// cloud's buffers)
template <int NORM>
__global__ void __launch_bounds__(kBwdThreads)
chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                   const int64_t* __restrict__ idx, const int64_t* __restrict__ len1,
                   const int64_t* __restrict__ len2, const float* __restrict__ weights, int P1, int P2,
                   int D, ChamferFeat ft, int reduction, int abs_cosine, int N,
                   const float* __restrict__ g_cham, const float* __restrict__ g_feat,
                   const int64_t* __restrict__ argmax, float* __restrict__ grad_x,
                   float* __restrict__ grad_y, int acc, int g_bcast, float g_scale) {
  const int n = blockIdx.y;
  int64_t L1l = len1[n], L2l = len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
  const bool y_empty = L2l <= 0;
  const float w = weights ? weights[n] : 1.0f;
  const float inv_len = 1.0f / static_cast<float>(L1 > 0 ? L1 : 1);
  const int64_t* in = idx + static_cast<size_t>(n) * P1;
  const float* xn = x + static_cast<size_t>(n) * P1 * D;
  const float* yn = y + static_cast<size_t>(n) * P2 * D;
  float* gxn = grad_x + static_cast<size_t>(n) * P1 * D;
  float* gyn = grad_y + static_cast<size_t>(n) * P2 * D;
  const int amax = (reduction == kRedMax) ? static_cast<int>(argmax[n]) : -1;

  for (int i = blockIdx.x * kBwdThreads + threadIdx.x; i < L1; i += gridDim.x * kBwdThreads) {
    // ---- chamfer term: d(dist)/dx, d(dist)/dy ----
    float gd;
    if (reduction == kRedNone) gd = g_cham[static_cast<size_t>(n) * P1 + i] * w;
    else if (reduction == kRedMax) gd = (i == amax) ? g_cham[n] * w : 0.0f;
    else gd = g_cham[g_bcast ? 0 : n] * g_scale * w * (reduction == kRedMean ? inv_len : 1.0f);  // g_bcast: one upstream scalar for every cloud
    int64_t j = in[i];
    const bool y_none = y_empty || j < 0 || j >= P2;  // out-of-range index: no neighbour, nothing read or scattered
    if (y_none) j = 0;
    if (gd != 0.0f && !y_none) {
      for (int d = 0; d < D; ++d) {
        const float a = xn[static_cast<size_t>(i) * D + d], b = yn[static_cast<size_t>(j) * D + d];
        const float diff = (NORM == 1) ? gd * ((a > b) ? 1.0f : -1.0f) : 2.0f * gd * (a - b);
        float* gp = gxn + static_cast<size_t>(i) * D + d;
        *gp = acc ? *gp + diff : diff;
        atomicAdd(gyn + static_cast<size_t>(j) * D + d, -diff);
      }
    }
    // ---- feature terms: d(1 - |cos|)/d(a), /d(b) ----
    for (int f = 0; f < ft.num; ++f) {
      const int C = ft.chans[f];
      float gf;
      if (reduction == kRedNone) gf = g_feat[(static_cast<size_t>(f) * N + n) * P1 + i] * w;
      else gf = g_feat[g_bcast ? static_cast<size_t>(f) : static_cast<size_t>(f) * N + n] * g_scale * w * (reduction == kRedMean ? inv_len : 1.0f);
      if (gf == 0.0f) continue;
      const float* a = ft.xf[f] + (static_cast<size_t>(n) * P1 + i) * C;
      const float* b = ft.yf[f] + (static_cast<size_t>(n) * P2 + j) * C;
      float na, nb;
      const float c = cosine(a, b, C, &na, &nb, y_none);
      float dc = -gf;  // d(1 - c)/dc
      if (abs_cosine) dc = (c > 0.0f) ? -gf : ((c < 0.0f) ? gf : 0.0f);
      // raw norms decide whether the clamp is active (then the norm is a constant)
      float sa = 0.f, sb = 0.f;
      for (int k = 0; k < C; ++k) {
        const float bv = y_none ? 0.0f : b[k];
        sa = fmaf(a[k], a[k], sa);
        sb = fmaf(bv, bv, sb);
      }
      const bool a_free = sqrtf(sa) > 1e-6f, b_free = sqrtf(sb) > 1e-6f;
      float* ga = ft.gxf[f] + (static_cast<size_t>(n) * P1 + i) * C;
      float* gb = ft.gyf[f] + (static_cast<size_t>(n) * P2 + j) * C;
      for (int k = 0; k < C; ++k) {
        const float bv = y_none ? 0.0f : b[k];
        const float ah = a[k] / na, bh = bv / nb;
        const float da = (bh - (a_free ? c * ah : 0.0f)) / na;
        const float db = (ah - (b_free ? c * bh : 0.0f)) / nb;
        ga[k] = acc ? ga[k] + dc * da : dc * da;
        if (!y_none) atomicAdd(gb + k, dc * db);
      }
    }
  }
}

}  // namespace pops

using namespace pops;

namespace {
int fill_feats(ChamferFeat* ft, int num_feats, const float* const* xf, const float* const* yf,
               float* const* gxf, float* const* gyf, const int64_t* chans) {
  if (num_feats < 0 || num_feats > kChamferMaxFeats) return 1;
  ft->num = num_feats;
  for (int f = 0; f < kChamferMaxFeats; ++f) {
    ft->xf[f] = f < num_feats ? xf[f] : nullptr;
    ft->yf[f] = f < num_feats ? yf[f] : nullptr;
    ft->gxf[f] = (f < num_feats && gxf) ? gxf[f] : nullptr;
    ft->gyf[f] = (f < num_feats && gyf) ? gyf[f] : nullptr;
    ft->chans[f] = f < num_feats ? static_cast<int>(chans[f]) : 0;
  }
  return 0;
}
}  // namespace

extern "C" int pops_chamfer_forward(const float* dists, const int64_t* idx, const int64_t* lengths1,
                                    const int64_t* lengths2, const float* weights, int64_t N,
                                    int64_t P1, int64_t P2, int num_feats, const float* const* xf,
                                    const float* const* yf, const int64_t* chans, int point_reduction,
                                    int abs_cosine, float* cham_out, float* feat_out,
                                    int64_t* argmax_out, pops_stream_t stream) {
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0, "negative size");
  POPS_CHECK_ARG(point_reduction >= 0 && point_reduction <= 3, "bad point_reduction");
  if (N == 0) return POPS_OK;
  POPS_CHECK_ARG(lengths1 && lengths2 && cham_out && (P1 == 0 || (dists && idx)), "null pointer argument");
  ChamferFeat ft;
  POPS_CHECK_ARG(fill_feats(&ft, num_feats, xf, yf, nullptr, nullptr, chans) == 0, "too many features (max 8)");
  POPS_CHECK_ARG(num_feats == 0 || feat_out, "null feat_out");
  POPS_CHECK_ARG(point_reduction != kRedMax || argmax_out, "null argmax_out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  profile_begin("chamfer", st);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kChamferCluster, static_cast<unsigned>(N));
    cfg.blockDim = dim3(kChamferThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kChamferCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, chamfer_fwd_kernel, dists, idx, lengths1, lengths2, weights, int(P1), int(P2),
                                    ft, point_reduction, abs_cosine, int(N), cham_out, feat_out, argmax_out));
  }
  profile_end("chamfer", st);
  POPS_LAUNCH_OK("chamfer_fwd_kernel");
  return POPS_OK;
}

extern "C" int pops_chamfer_backward(const float* x, const float* y, const int64_t* idx,
                                     const int64_t* lengths1, const int64_t* lengths2,
                                     const float* weights, int64_t N, int64_t P1, int64_t P2, int64_t D,
                                     int norm, int num_feats, const float* const* xf,
                                     const float* const* yf, const int64_t* chans, int point_reduction,
                                     int abs_cosine, const float* g_cham, const float* g_feat,
                                     const int64_t* argmax, float* grad_x, float* grad_y,
                                     float* const* grad_xf, float* const* grad_yf, int accumulate,
                                     int g_broadcast, float g_scale, pops_stream_t stream) {
  POPS_CHECK_ARG(norm == 1 || norm == 2, "Norm must be 1 or 2.");
  POPS_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 0, "negative size");
  POPS_CHECK_ARG(point_reduction >= 0 && point_reduction <= 3, "bad point_reduction");
  POPS_CHECK_ARG(!g_broadcast || point_reduction == kRedSum || point_reduction == kRedMean,
                 "g_broadcast needs point_reduction sum or mean");
  ChamferFeat ft;
  POPS_CHECK_ARG(fill_feats(&ft, num_feats, xf, yf, grad_xf, grad_yf, chans) == 0, "too many features (max 8)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate) {
    if (N * P1 * D > 0) POPS_CUDA_OK(cudaMemsetAsync(grad_x, 0, size_t(N) * P1 * D * 4, st));
    if (N * P2 * D > 0) POPS_CUDA_OK(cudaMemsetAsync(grad_y, 0, size_t(N) * P2 * D * 4, st));
    for (int f = 0; f < num_feats; ++f) {
      if (N * P1 * chans[f] > 0) POPS_CUDA_OK(cudaMemsetAsync(ft.gxf[f], 0, size_t(N) * P1 * chans[f] * 4, st));
      if (N * P2 * chans[f] > 0) POPS_CUDA_OK(cudaMemsetAsync(ft.gyf[f], 0, size_t(N) * P2 * chans[f] * 4, st));
    }
  }
  if (N == 0 || P1 == 0) return POPS_OK;
  POPS_CHECK_ARG(x && y && idx && lengths1 && lengths2 && g_cham, "null pointer argument");
  const dim3 bgrid(static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(P1, kBwdThreads), 64))),
                   static_cast<unsigned>(N));
  if (norm == 2)
    chamfer_bwd_kernel<2><<<bgrid, kBwdThreads, 0, st>>>(
        x, y, idx, lengths1, lengths2, weights, int(P1), int(P2), int(D), ft, point_reduction, abs_cosine,
        int(N), g_cham, g_feat, argmax, grad_x, grad_y, accumulate ? 1 : 0, g_broadcast ? 1 : 0, g_scale);
  else
    chamfer_bwd_kernel<1><<<bgrid, kBwdThreads, 0, st>>>(
        x, y, idx, lengths1, lengths2, weights, int(P1), int(P2), int(D), ft, point_reduction, abs_cosine,
        int(N), g_cham, g_feat, argmax, grad_x, grad_y, accumulate ? 1 : 0, g_broadcast ? 1 : 0, g_scale);
  POPS_LAUNCH_OK("chamfer_bwd_kernel");
  return POPS_OK;
}
