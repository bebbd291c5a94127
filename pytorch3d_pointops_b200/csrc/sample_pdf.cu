// sample_pdf for sm_100a: inverse-CDF sampling of piecewise-constant densities.
//
// Replaces csrc/sample_pdf/{sample_pdf.cu,sample_pdf_cpu.cpp}.  Contract = the reference CPU path
// (sample_pdf_cpu.cpp:24-100, which defines USE_BINARY_SEARCH): per row the weights are summed
// left to right in float32 (running sums kept), a quantile u read from `outputs` becomes
// q = fl((total + eps) * u), its bin is the first of the first n_bins-1 running sums that is not
// below q, and the sample is bin_start + (q' / w) * (bin_end - bin_start) with the same float
// operations (no FMA: the library is built with -fmad=false; division is IEEE).  `outputs` is
// overwritten in place.
//
// One CTA per row: thread 0 forms the running sums in the reference's order (a few hundred
// dependent adds), everybody then takes samples with a shared-memory binary search.  HBM traffic =
// bins + weights once, outputs read + written once.
#include "common.cuh"

namespace pops {

constexpr int kPdfThreads = 128;

__global__ void __launch_bounds__(kPdfThreads)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, float* __restrict__ outputs,
                  int n_bins, int n_samples, float eps) {
  extern __shared__ float sm[];  // partial[n_bins], w[n_bins], edge[n_bins + 1]
  float* partial = sm;
  float* w = sm + n_bins;
  float* edge = w + n_bins;
  __shared__ float s_total;
  const size_t b = blockIdx.x;
  const float* wb = weights + b * n_bins;
  const float* eb = bins + b * (n_bins + 1);
  for (int i = threadIdx.x; i < n_bins; i += kPdfThreads) w[i] = wb[i];
  for (int i = threadIdx.x; i <= n_bins; i += kPdfThreads) edge[i] = eb[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.0f;
    for (int i = 0; i < n_bins; ++i) {
      total = __fadd_rn(total, w[i]);
      partial[i] = total;
    }
    s_total = __fadd_rn(total, eps);
  }
  __syncthreads();
  const float total = s_total;
  float* out = outputs + b * n_samples;
  for (int s = threadIdx.x; s < n_samples; s += kPdfThreads) {
    float q = __fmul_rn(total, out[s]);
    int lo = 0, hi = n_bins - 1;  // std::lower_bound over partial[0 .. n_bins-1)
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);
      if (partial[mid] < q) lo = mid + 1; else hi = mid;
    }
    if (lo > 0) q = __fsub_rn(q, partial[lo - 1]);
    const float b0 = edge[lo], b1 = edge[lo + 1], bw = w[lo];
    float v = b0;
    if (q > bw) {
      v = b1;
    } else if (bw > eps) {
      v = __fadd_rn(b0, __fmul_rn(__fdiv_rn(q, bw), __fsub_rn(b1, b0)));
    }
    out[s] = v;
  }
}

}  // namespace pops

using namespace pops;

extern "C" int pops_sample_pdf(const float* bins, const float* weights, float* outputs, int64_t B,
                               int64_t n_bins, int64_t n_samples, float eps, pops_stream_t stream) {
  POPS_CHECK_ARG(B >= 0 && n_bins >= 1 && n_samples >= 0, "bad sizes");
  if (B == 0 || n_samples == 0) return POPS_OK;
  POPS_CHECK_ARG(bins && weights && outputs, "null pointer argument");
  POPS_CHECK_ARG(B < (int64_t(1) << 31) && n_samples < (int64_t(1) << 31), "size too large");
  const size_t smem = (size_t(3) * n_bins + 1) * 4;
  if (smem > 200 * 1024) return fail(POPS_ERR_UNSUPPORTED, "sample_pdf: too many bins");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  POPS_CUDA_OK(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  sample_pdf_kernel<<<static_cast<unsigned>(B), kPdfThreads, smem, st>>>(bins, weights, outputs, int(n_bins),
                                                                         int(n_samples), eps);
  POPS_LAUNCH_OK("sample_pdf_kernel");
  return POPS_OK;
}
