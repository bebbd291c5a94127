// Ragged copies (packed <-> padded) and fused row gathers for sm_100a.  All HBM-bound.
//
// packed_to_padded / padded_to_packed replace csrc/packed_to_padded_tensor/{.cu,_cpu.cpp}
// (packed_to_padded_tensor_cpu.cpp:11-70).  The reference launches one 512-thread block per
// cloud (packed_to_padded_tensor.cu:149-150) -- load-imbalanced for ragged batches; here the
// grid is flat over OUTPUT elements, each thread resolves its cloud from first_idxs, so the
// work is balanced and every output element (zero padding included) is written exactly once.
//
// gather replaces the torch expand+gather+mask sequences of knn_gather (functions/knn.py:200-250)
// and masked_gather (functions/utils.py:20-65): one pass, idx read once (8 B), row read + written.
#include "common.cuh"

namespace pops {

// padded[b, i, :] = i < num_b ? packed[first[b] + i, :] : 0
__global__ void packed_to_padded_kernel(const float* __restrict__ packed,
                                        const int64_t* __restrict__ first, int64_t num_inputs,
                                        int B, int64_t max_size, int D, float* __restrict__ padded) {
  const int64_t total = static_cast<int64_t>(B) * max_size * D;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(e % D);
    const int64_t row = e / D;
    const int b = static_cast<int>(row / max_size);
    const int64_t i = row % max_size;
    const int64_t start = first[b];
    const int64_t end = (b + 1 < B) ? first[b + 1] : num_inputs;
    float v = 0.0f;
    if (i < end - start && start + i < num_inputs && start + i >= 0) v = packed[(start + i) * D + d];
    padded[e] = v;
  }
}

// packed[f, :] = padded[b(f), f - first[b], :]  (0 when the row is not covered by any cloud or
// lies beyond max_size).  b(f) by binary search over first_idxs (non-decreasing).
__global__ void padded_to_packed_kernel(const float* __restrict__ padded,
                                        const int64_t* __restrict__ first, int64_t num_inputs,
                                        int B, int64_t max_size, int D, float* __restrict__ packed) {
  const int64_t total = num_inputs * D;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(e % D);
    const int64_t f = e / D;
    // last b with first[b] <= f  (for equal starts the LAST cloud is the non-empty one)
    int lo = 0, hi = B;  // invariant: first[lo] <= f (if any), answer in [lo, hi)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (first[mid] <= f) lo = mid; else hi = mid;
    }
    float v = 0.0f;
    const int64_t start = first[lo];
    const int64_t i = f - start;
    if (i >= 0 && i < max_size) v = padded[(static_cast<int64_t>(lo) * max_size + i) * D + d];
    packed[e] = v;
  }
}

// out[n, l, k, :] = x[n, idx[n,l,k], :] with masking.  One thread per (row, 4-float chunk) when
// U % 4 == 0 and rows are 16-byte aligned, else one thread per element.
template <int MODE, int VEC>
__global__ void gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                              const int64_t* __restrict__ lengths, int64_t rows_per_cloud /*L*K*/,
                              int K, int M, int U, int64_t total_rows, float* __restrict__ out,
                              int32_t* __restrict__ oob) {
  const int UV = U / VEC;
  const int64_t total = total_rows * UV;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(e % UV);
    const int64_t row = e / UV;  // (n*L + l)*K + k
    const int n = static_cast<int>(row / rows_per_cloud);
    const int k = static_cast<int>(row % K);
    int64_t j = idx[row];
    bool take = true;
    if (MODE == POPS_GATHER_KNN) {
      if (lengths != nullptr && k >= lengths[n]) take = false;
      if (take && (j < 0 || j >= M)) {
        take = false;
        if (oob != nullptr) *oob = 1;
      }
    } else {
      if (j < 0 || j >= M) take = false;  // -1 = padding
    }
    const float* src = x + (static_cast<int64_t>(n) * M + (take ? j : 0)) * U;
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (take) v = reinterpret_cast<const float4*>(src)[u];
      reinterpret_cast<float4*>(out + row * U)[u] = v;
    } else {
      out[row * U + u] = take ? src[u] : 0.0f;
    }
  }
}

template <int MODE>
__global__ void gather_backward_kernel(const float* __restrict__ grad_out,
                                       const int64_t* __restrict__ idx,
                                       const int64_t* __restrict__ lengths,
                                       int64_t rows_per_cloud, int K, int M, int U,
                                       int64_t total_rows, float* __restrict__ grad_x) {
  const int64_t total = total_rows * U;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(e % U);
    const int64_t row = e / U;
    const int n = static_cast<int>(row / rows_per_cloud);
    const int k = static_cast<int>(row % K);
    const int64_t j = idx[row];
    if (j < 0 || j >= M) continue;
    if (MODE == POPS_GATHER_KNN && lengths != nullptr && k >= lengths[n]) continue;
    atomicAdd(grad_x + (static_cast<int64_t>(n) * M + j) * U + u, grad_out[e]);
  }
}

// Fused neighbourhood covariance (functions/utils.py:111-153 of the reference): per query row the K
// gathered neighbours nn[k] = x[n, idx[k]] (zero where k >= lengths[n], as knn_gather does), their
// mean over ALL K slots, and cov = mean_k (nn[k] - mean)(nn[k] - mean)^T.  One thread per query row
// writes nn (N,P,K,DT) and cov (N,P,DT,DT); the reference materialises (N,P,K,D) centred values and
// (N,P,K,D,D) outer products in between.
template <int DT>
__global__ void point_cov_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                                 const int64_t* __restrict__ lengths, int P, int M, int K,
                                 float* __restrict__ nn, float* __restrict__ cov) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t row = static_cast<size_t>(n) * P + p;
  const int64_t len = lengths ? lengths[n] : K;
  const float* xn = x + static_cast<size_t>(n) * M * DT;
  const int64_t* ir = idx + row * K;
  float* nr = nn + row * K * DT;
  float mean[DT];
#pragma unroll
  for (int d = 0; d < DT; ++d) mean[d] = 0.0f;
  for (int k = 0; k < K; ++k) {
    const int64_t j = ir[k];
    const bool take = k < len && j >= 0 && j < M;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      const float v = take ? xn[j * DT + d] : 0.0f;
      nr[k * DT + d] = v;
      mean[d] += v;
    }
  }
  const float invK = 1.0f / static_cast<float>(K);
#pragma unroll
  for (int d = 0; d < DT; ++d) mean[d] *= invK;
  float c[DT][DT];
#pragma unroll
  for (int a = 0; a < DT; ++a)
#pragma unroll
    for (int b = 0; b < DT; ++b) c[a][b] = 0.0f;
  for (int k = 0; k < K; ++k) {
    float v[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = nr[k * DT + d] - mean[d];  // re-read of what this thread just wrote
#pragma unroll
    for (int a = 0; a < DT; ++a)
#pragma unroll
      for (int b = 0; b < DT; ++b) c[a][b] += v[a] * v[b];
  }
  float* cr = cov + row * DT * DT;
#pragma unroll
  for (int a = 0; a < DT; ++a)
#pragma unroll
    for (int b = 0; b < DT; ++b) cr[a * DT + b] = c[a][b] * invK;
}

inline int flat_grid(int64_t total, int threads) {
  return int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), int64_t(num_sms()) * 32)));
}

}  // namespace pops

using namespace pops;

extern "C" int pops_packed_to_padded(const float* packed, const int64_t* first_idxs,
                                     int64_t num_inputs, int64_t B, int64_t max_size, int64_t D,
                                     float* padded, pops_stream_t stream) {
  POPS_CHECK_ARG(num_inputs >= 0 && B >= 0 && max_size >= 0 && D >= 0, "negative size");
  const int64_t total = B * max_size * D;
  if (total == 0) return POPS_OK;
  POPS_CHECK_ARG(first_idxs && padded && (packed || num_inputs == 0), "null pointer argument");
  POPS_CHECK_ARG(B < (int64_t(1) << 31) && D < (int64_t(1) << 31), "size too large");
  packed_to_padded_kernel<<<flat_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, first_idxs, num_inputs, int(B), max_size, int(D), padded);
  POPS_LAUNCH_OK("packed_to_padded_kernel");
  return POPS_OK;
}

extern "C" int pops_padded_to_packed(const float* padded, const int64_t* first_idxs,
                                     int64_t num_inputs, int64_t B, int64_t max_size, int64_t D,
                                     float* packed, pops_stream_t stream) {
  POPS_CHECK_ARG(num_inputs >= 0 && B >= 0 && max_size >= 0 && D >= 0, "negative size");
  const int64_t total = num_inputs * D;
  if (total == 0) return POPS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  POPS_CHECK_ARG(packed, "null pointer argument");
  if (B == 0 || max_size == 0) {
    POPS_CUDA_OK(cudaMemsetAsync(packed, 0, size_t(total) * 4, st));
    return POPS_OK;
  }
  POPS_CHECK_ARG(first_idxs && padded, "null pointer argument");
  POPS_CHECK_ARG(B < (int64_t(1) << 31) && D < (int64_t(1) << 31), "size too large");
  padded_to_packed_kernel<<<flat_grid(total, 256), 256, 0, st>>>(padded, first_idxs, num_inputs,
                                                                 int(B), max_size, int(D), packed);
  POPS_LAUNCH_OK("padded_to_packed_kernel");
  return POPS_OK;
}

extern "C" int pops_gather(const float* x, const int64_t* idx, const int64_t* lengths, int64_t N,
                           int64_t M, int64_t U, int64_t L, int64_t K, int mode, float* out,
                           int32_t* oob_flag, pops_stream_t stream) {
  POPS_CHECK_ARG(mode == POPS_GATHER_KNN || mode == POPS_GATHER_MASKED, "bad gather mode");
  POPS_CHECK_ARG(N >= 0 && M >= 0 && U >= 0 && L >= 0 && K >= 0, "negative size");
  const int64_t rows = N * L * K;
  if (rows * U == 0) return POPS_OK;
  POPS_CHECK_ARG(idx && out && (x || M == 0), "null pointer argument");
  POPS_CHECK_ARG(M < (int64_t(1) << 31) && U < (int64_t(1) << 31) && K < (int64_t(1) << 31), "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (U % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int64_t total = rows * (vec ? U / 4 : U);
  const int grid = flat_grid(total, 256);
#define POPS_GATHER(MODE, VEC)                                                                  \
  gather_kernel<MODE, VEC><<<grid, 256, 0, st>>>(x, idx, lengths, L * K, int(K), int(M), int(U), \
                                                 rows, out, oob_flag)
  if (mode == POPS_GATHER_KNN) { if (vec) POPS_GATHER(POPS_GATHER_KNN, 4); else POPS_GATHER(POPS_GATHER_KNN, 1); }
  else { if (vec) POPS_GATHER(POPS_GATHER_MASKED, 4); else POPS_GATHER(POPS_GATHER_MASKED, 1); }
#undef POPS_GATHER
  POPS_LAUNCH_OK("gather_kernel");
  return POPS_OK;
}

extern "C" int pops_gather_backward(const float* grad_out, const int64_t* idx,
                                    const int64_t* lengths, int64_t N, int64_t M, int64_t U,
                                    int64_t L, int64_t K, int mode, float* grad_x,
                                    pops_stream_t stream) {
  POPS_CHECK_ARG(mode == POPS_GATHER_KNN || mode == POPS_GATHER_MASKED, "bad gather mode");
  POPS_CHECK_ARG(N >= 0 && M >= 0 && U >= 0 && L >= 0 && K >= 0, "negative size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N * M * U > 0) {
    POPS_CHECK_ARG(grad_x, "null pointer argument");
    POPS_CUDA_OK(cudaMemsetAsync(grad_x, 0, size_t(N) * M * U * 4, st));
  }
  const int64_t rows = N * L * K;
  if (rows * U == 0 || M == 0) return POPS_OK;
  POPS_CHECK_ARG(grad_out && idx, "null pointer argument");
  const int grid = flat_grid(rows * U, 256);
  if (mode == POPS_GATHER_KNN)
    gather_backward_kernel<POPS_GATHER_KNN><<<grid, 256, 0, st>>>(grad_out, idx, lengths, L * K, int(K), int(M), int(U), rows, grad_x);
  else
    gather_backward_kernel<POPS_GATHER_MASKED><<<grid, 256, 0, st>>>(grad_out, idx, lengths, L * K, int(K), int(M), int(U), rows, grad_x);
  POPS_LAUNCH_OK("gather_backward_kernel");
  return POPS_OK;
}

extern "C" int pops_point_covariances(const float* x, const int64_t* idx, const int64_t* lengths, int64_t N,
                                      int64_t P, int64_t M, int64_t D, int64_t K, float* nn, float* cov,
                                      pops_stream_t stream) {
  POPS_CHECK_ARG(N >= 0 && P >= 0 && M >= 0 && K >= 1, "bad sizes");
  if (D < 1 || D > 4) return fail(POPS_ERR_UNSUPPORTED, "point_covariances: fused path covers 1 <= D <= 4");
  if (N == 0 || P == 0) return POPS_OK;
  POPS_CHECK_ARG(x && idx && nn && cov, "null pointer argument");
  POPS_CHECK_ARG(N < 65536 && P < (int64_t(1) << 31) && M < (int64_t(1) << 31), "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(ceil_div(P, 128)), static_cast<unsigned>(N));
  switch (D) {
    case 1: point_cov_kernel<1><<<grid, 128, 0, st>>>(x, idx, lengths, int(P), int(M), int(K), nn, cov); break;
    case 2: point_cov_kernel<2><<<grid, 128, 0, st>>>(x, idx, lengths, int(P), int(M), int(K), nn, cov); break;
    case 3: point_cov_kernel<3><<<grid, 128, 0, st>>>(x, idx, lengths, int(P), int(M), int(K), nn, cov); break;
    default: point_cov_kernel<4><<<grid, 128, 0, st>>>(x, idx, lengths, int(P), int(M), int(K), nn, cov); break;
  }
  POPS_LAUNCH_OK("point_cov_kernel");
  return POPS_OK;
}
