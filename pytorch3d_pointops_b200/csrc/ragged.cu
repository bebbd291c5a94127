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

// Both copies are batches of contiguous segment copies: cloud b owns the packed floats
// [first[b]*D, first[b+1]*D) and the padded floats [b*max*D, (b+1)*max*D).  grid.y walks the clouds,
// so no thread divides by D or max_size; the destination is written as aligned 16-byte chunks
// (chunk c of the cloud's segment covers floats [4c - a, 4c - a + 4), a = misalignment of the
// segment start in floats), the source with 4-byte loads (its alignment differs from the
// destination's; the four loads of a chunk hit the same L1 lines).  Every destination float is
// written exactly once, zero padding included.
constexpr int kCopyUnroll = 4;

__device__ __forceinline__ void store_chunk(float* __restrict__ dst_seg, int64_t f0, int64_t F, const float (&v)[4]) {
  if (f0 >= 0 && f0 + 4 <= F) {
    *reinterpret_cast<float4*>(dst_seg + f0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (f0 + e >= 0 && f0 + e < F) dst_seg[f0 + e] = v[e];
  }
}

// padded[b, i, :] = i < num_b ? packed[first[b] + i, :] : 0
__global__ void __launch_bounds__(256)
packed_to_padded_kernel(const float* __restrict__ packed, const int64_t* __restrict__ first,
                        int64_t num_inputs, int B, int64_t max_size, int D, float* __restrict__ padded) {
  const int64_t F = max_size * D;  // floats per padded cloud
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const int64_t start = first[b];
    const int64_t end = (b + 1 < B) ? first[b + 1] : num_inputs;
    // rows [start, start + num) exist in packed; clamp against garbage first_idxs
    int64_t num = end - start;
    if (start < 0 || num < 0) num = 0;
    if (num > max_size) num = max_size;
    if (start + num > num_inputs) num = num_inputs - start > 0 ? num_inputs - start : 0;
    const int64_t live = num * D;  // floats copied; the rest of the segment is zero
    float* dst = padded + static_cast<int64_t>(b) * F;
    const float* src = packed + start * D;
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
    const int64_t nchunks = (F + a + 3) >> 2;
    // (one chunk per thread: unrolling four, as padded_to_packed_kernel does, measured SLOWER here -- 108 vs
    //  96 us on 64 clouds x 65536 rows x 16 floats, 26.6 vs 24.6 us at D = 3.  Two aligned 16-byte loads and
    //  a funnel select per chunk instead of four 4-byte loads were slower too, in both kernels: twice the L1
    //  bytes -- D = 3 24.6 -> 28.7 us here, 22.5 -> 41 us for padded_to_packed.  An unpredicated fast path for
    //  interior chunks -- one address, four 4-byte loads, one store -- changed nothing: the four L1 wavefronts
    //  per 16 bytes are the limit, not the issue slots.  One aligned 16-byte load per lane with the missing
    //  floats taken from the next lane by shuffle was slower again: 34.8 us at D = 3, 141 us at D = 16.)
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < nchunks;
         c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t f0 = 4 * c - a;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t f = f0 + e;
        v[e] = (f >= 0 && f < live) ? __ldg(src + f) : 0.0f;
      }
      store_chunk(dst, f0, F, v);
    }
  }
}

// packed[f, :] = padded[b(f), f - first[b], :] for first[b] <= f < first[b+1], 0 beyond max_size and
// for rows no cloud covers (f < first[0]); cloud 0 also zero-fills the rows before first[0].
__global__ void __launch_bounds__(256)
padded_to_packed_kernel(const float* __restrict__ padded, const int64_t* __restrict__ first,
                        int64_t num_inputs, int B, int64_t max_size, int D, float* __restrict__ packed) {
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    int64_t start = first[b];
    int64_t end = (b + 1 < B) ? first[b + 1] : num_inputs;
    start = start < 0 ? 0 : (start > num_inputs ? num_inputs : start);
    end = end < start ? start : (end > num_inputs ? num_inputs : end);
    int64_t num = end - start;
    if (num > max_size) num = max_size;
    const int64_t live = num * D;
    const int64_t lead = (b == 0) ? start * D : 0;  // zero rows in front of the first cloud
    const int64_t F = (end - start) * D + lead;
    float* dst = packed + start * D - lead;
    const float* src = padded + static_cast<int64_t>(b) * max_size * D;
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
    const int64_t nchunks = (F + a + 3) >> 2;
    // kCopyUnroll chunks per thread and trip, a block apart (coalesced): all 16 loads are requested before
    // the first store (64 clouds x 65536 rows: D = 3 25.6 -> 22.5 us, D = 16 100 -> 80 us)
    for (int64_t c0 = static_cast<int64_t>(blockIdx.x) * blockDim.x * kCopyUnroll + threadIdx.x; c0 < nchunks;
         c0 += static_cast<int64_t>(gridDim.x) * blockDim.x * kCopyUnroll) {
      float v[kCopyUnroll][4];
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) {
        const int64_t f0 = 4 * (c0 + k * blockDim.x) - a;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t f = f0 + e - lead;
          v[k][e] = (f >= 0 && f < live) ? __ldg(src + f) : 0.0f;
        }
      }
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) {
        const int64_t c = c0 + k * blockDim.x;
        if (c < nchunks) store_chunk(dst, 4 * c - a, F, v[k]);
      }
    }
  }
}

// out[n, l, k, :] = x[n, idx[n,l,k], :] with masking.  grid.y walks the clouds; within a cloud the
// OUTPUT is a flat float stream written as aligned 16-byte chunks (perfectly coalesced stores for
// any U, e.g. the 12-byte rows of U = 3); a chunk spans at most RMAX rows.  The kernel is a chain of
// two dependent loads (index from HBM, then the row from L1/L2), so every thread works on UN chunks
// at once: all their indices are requested first, then all their rows, then the stores -- without
// that the SMs sit at full occupancy waiting (measured: 0.26 of HBM with one chunk per thread).
// UT: compile-time U (0 = runtime).  V4: U % 4 == 0 and 16-byte aligned x -> one 16-byte source
// load per chunk.
template <int MODE, int UT, bool V4>
__global__ void __launch_bounds__(256)
gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
              const int64_t* __restrict__ lengths, unsigned LK /*L*K rows per cloud*/, unsigned K, int M,
              unsigned U_rt, int N, float* __restrict__ out, int32_t* __restrict__ oob) {
  constexpr int UN = 4;
  constexpr int RMAX = V4 ? 1 : (UT == 0 || UT == 1 ? 4 : (UT == 2 ? 3 : 2));  // rows a 4-float chunk can touch
  const unsigned U = UT ? UT : U_rt;
  const int64_t F = static_cast<int64_t>(LK) * U;  // floats of one cloud's output (< 2^31, host-checked)
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    float* dst = out + static_cast<int64_t>(n) * F;
    const int64_t* idx_n = idx + static_cast<int64_t>(n) * LK;
    const float* x_n = x + static_cast<int64_t>(n) * M * U;
    // knn_gather zeroes slots k >= lengths[n]; only clouds shorter than K need k at all
    unsigned klim = K;
    if (MODE == POPS_GATHER_KNN && lengths != nullptr) {
      const int64_t len = lengths[n];
      klim = len < 0 ? 0u : (len < static_cast<int64_t>(K) ? static_cast<unsigned>(len) : K);
    }
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
    const unsigned nchunks = static_cast<unsigned>((F + a + 3) >> 2);
    const unsigned step = gridDim.x * blockDim.x * UN;
    for (unsigned cb = blockIdx.x * blockDim.x * UN + threadIdx.x; cb < nchunks; cb += step) {
      int f0[UN];
      unsigned row[UN], u[UN];
      long long jraw[UN][RMAX];
      // ---- all indices first: unconditional loads from clamped addresses, no branch between them ----
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const unsigned c = cb + k * blockDim.x;
        f0[k] = static_cast<int>(4u * c) - a;
        const unsigned fs = f0[k] < 0 ? 0u : static_cast<unsigned>(f0[k]);
        row[k] = fs / U;
        u[k] = fs - row[k] * U;
#pragma unroll
        for (int i = 0; i < RMAX; ++i) jraw[k][i] = __ldg(idx_n + min(row[k] + i, LK - 1u));
      }
      int jv[UN][RMAX];  // source row of the chunk's i-th output row, -1 = zeros
      bool bad = false;
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const unsigned c = cb + k * blockDim.x;
        const unsigned span = u[k] + 4u - (f0[k] < 0 ? static_cast<unsigned>(-f0[k]) : 0u);  // floats from the start of row[k] to the chunk's end
#pragma unroll
        for (int i = 0; i < RMAX; ++i) {
          const unsigned r = row[k] + i;
          const long long j = jraw[k][i];
          const bool need = c < nchunks && r < LK && static_cast<unsigned>(i) * U < span;
          bool take = need && j >= 0 && j < M;
          if (MODE == POPS_GATHER_KNN) {
            const bool live = klim == K || (r % K) < klim;
            bad = bad || (need && live && !take);
            take = take && live;
          }
          jv[k][i] = take ? static_cast<int>(j) : -1;
        }
      }
      if (MODE == POPS_GATHER_KNN && bad && oob != nullptr) *oob = 1;
      // ---- then all rows (row 0 of the cloud stands in for "zeros": always a valid address) ----
      float v[UN][4];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        if (V4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(x_n + static_cast<int64_t>(max(jv[k][0], 0)) * U + u[k]));
          const bool t = jv[k][0] >= 0;
          v[k][0] = t ? s4.x : 0.0f; v[k][1] = t ? s4.y : 0.0f; v[k][2] = t ? s4.z : 0.0f; v[k][3] = t ? s4.w : 0.0f;
        } else {
          const int lead = f0[k] < 0 ? -f0[k] : 0;  // floats of this chunk in front of the cloud's segment
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            unsigned off = u[k] + static_cast<unsigned>(e >= lead ? e - lead : 0);  // float offset from the start of row[k]
            int i = 0;
#pragma unroll
            for (int t = 1; t < RMAX; ++t)
              if (off >= U) { off -= U; i = t; }
            int j = jv[k][0];
#pragma unroll
            for (int t = 1; t < RMAX; ++t) j = (i == t) ? jv[k][t] : j;
            const bool t = e >= lead && j >= 0 && off < U;
            const float val = __ldg(x_n + static_cast<int64_t>(t ? j : 0) * U + (t ? off : 0u));
            v[k][e] = t ? val : 0.0f;
          }
        }
      }
      // ---- then the stores ----
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (cb + k * blockDim.x < nchunks) store_chunk(dst, f0[k], F, v[k]);
    }
  }
}

// U = 3 (xyz rows, the shape of every knn_gather / masked_gather on points) with L*K % 4 == 0 and
// 16-byte aligned buffers: one thread owns FOUR consecutive output rows = 48 bytes = three aligned
// 16-byte stores, reads their four indices as two 16-byte loads, and fetches every 12-byte source row
// with one 8-byte and one 4-byte load (which of the two comes first depends on the row's parity).
// The gather is bound by L1 sector lookups, not by HBM (ncu: l1tex 63 %, dram 21 % for the generic
// kernel): this form needs 2.5 lookups per row instead of 4.5.
template <int MODE>
__global__ void __launch_bounds__(256)
gather_rows3_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                    const int64_t* __restrict__ lengths, unsigned LK, unsigned K, int M, int N,
                    float* __restrict__ out, int32_t* __restrict__ oob) {
  const unsigned quads = LK >> 2;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    float* dst = out + static_cast<int64_t>(n) * LK * 3;
    const int64_t* idx_n = idx + static_cast<int64_t>(n) * LK;
    const float* x_n = x + static_cast<int64_t>(n) * M * 3;
    unsigned klim = K;
    if (MODE == POPS_GATHER_KNN && lengths != nullptr) {
      const int64_t len = lengths[n];
      klim = len < 0 ? 0u : (len < static_cast<int64_t>(K) ? static_cast<unsigned>(len) : K);
    }
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < quads; t += gridDim.x * blockDim.x) {
      const unsigned r0 = t << 2;
      const longlong2 ja = __ldg(reinterpret_cast<const longlong2*>(idx_n + r0));
      const longlong2 jb = __ldg(reinterpret_cast<const longlong2*>(idx_n + r0 + 2));
      const long long j[4] = {ja.x, ja.y, jb.x, jb.y};
      bool take[4];
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        take[i] = j[i] >= 0 && j[i] < M;
        if (MODE == POPS_GATHER_KNN) {
          const bool live = klim == K || ((r0 + i) % K) < klim;
          bad = bad || (live && !take[i]);
          take[i] = take[i] && live;
        }
      }
      if (MODE == POPS_GATHER_KNN && bad && oob != nullptr) *oob = 1;
      float2 a[4];
      float b[4];
      bool al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // all eight loads first (row 0 of the cloud stands in for "zeros")
        const float* row = x_n + (take[i] ? j[i] : 0) * 3;
        al[i] = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
        a[i] = __ldg(reinterpret_cast<const float2*>(al[i] ? row : row + 1));
        b[i] = __ldg(al[i] ? row + 2 : row);
      }
      float v[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[3 * i + 0] = take[i] ? (al[i] ? a[i].x : b[i]) : 0.0f;
        v[3 * i + 1] = take[i] ? (al[i] ? a[i].y : a[i].x) : 0.0f;
        v[3 * i + 2] = take[i] ? (al[i] ? b[i] : a[i].y) : 0.0f;
      }
      float4* o = reinterpret_cast<float4*>(dst + static_cast<int64_t>(r0) * 3);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
      o[2] = make_float4(v[8], v[9], v[10], v[11]);
    }
  }
}

// U = 3, a cloud that fits shared memory (M * 12 bytes <= ~220 KB: 16384-point clouds), many more gathered
// rows than points: the CTA copies ITS cloud into shared memory once (coalesced 16-byte loads, L2 hits
// after the first CTA of the cloud) and every source row is then a shared-memory read.  What is left
// for the memory system is the stream itself -- indices in (8 B per row), rows out (12 B per row) -- so
// the kernel runs at the HBM rate instead of the L1/L2 sector-lookup rate of gather_rows3_kernel.
// grid (ctas per cloud, N), 1024 threads, one CTA per SM; same row ownership (4 consecutive rows per
// thread: two 16-byte index loads, three 16-byte stores) and the same masking as gather_rows3_kernel.
constexpr int kGatherSmemThreads = 1024;

template <int MODE>
__global__ void __launch_bounds__(kGatherSmemThreads, 1)
gather_rows3_smem_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                         const int64_t* __restrict__ lengths, unsigned LK, unsigned K, int M, int N,
                         float* __restrict__ out, int32_t* __restrict__ oob) {
  extern __shared__ __align__(16) float cloud[];  // [M][3]
  const unsigned quads = LK >> 2;
  const int n = blockIdx.y;
  float* dst = out + static_cast<int64_t>(n) * LK * 3;
  const int64_t* idx_n = idx + static_cast<int64_t>(n) * LK;
  const float* x_n = x + static_cast<int64_t>(n) * M * 3;
  {
    const int nf4 = (M * 3) >> 2;  // (M * 3) % 4 == 0 and x_n 16-byte aligned: host-checked
    const float4* src = reinterpret_cast<const float4*>(x_n);
    float4* d4 = reinterpret_cast<float4*>(cloud);
    for (int i = threadIdx.x; i < nf4; i += kGatherSmemThreads) d4[i] = __ldg(src + i);
  }
  unsigned klim = K;
  if (MODE == POPS_GATHER_KNN && lengths != nullptr) {
    const int64_t len = lengths[n];
    klim = len < 0 ? 0u : (len < static_cast<int64_t>(K) ? static_cast<unsigned>(len) : K);
  }
  __syncthreads();
  for (unsigned t = blockIdx.x * kGatherSmemThreads + threadIdx.x; t < quads; t += gridDim.x * kGatherSmemThreads) {
    const unsigned r0 = t << 2;
    const longlong2 ja = __ldcs(reinterpret_cast<const longlong2*>(idx_n + r0));
    const longlong2 jb = __ldcs(reinterpret_cast<const longlong2*>(idx_n + r0 + 2));
    const long long j[4] = {ja.x, ja.y, jb.x, jb.y};
    bool bad = false;
    float v[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool take = j[i] >= 0 && j[i] < M;
      if (MODE == POPS_GATHER_KNN) {
        const bool live = klim == K || ((r0 + i) % K) < klim;
        bad = bad || (live && !take);
        take = take && live;
      }
      const float* row = cloud + (take ? static_cast<int>(j[i]) : 0) * 3;
      const float a = row[0], b = row[1], c = row[2];
      v[3 * i + 0] = take ? a : 0.0f;
      v[3 * i + 1] = take ? b : 0.0f;
      v[3 * i + 2] = take ? c : 0.0f;
    }
    if (MODE == POPS_GATHER_KNN && bad && oob != nullptr) *oob = 1;
    float4* o = reinterpret_cast<float4*>(dst + static_cast<int64_t>(r0) * 3);
    __stcs(o, make_float4(v[0], v[1], v[2], v[3]));
    __stcs(o + 1, make_float4(v[4], v[5], v[6], v[7]));
    __stcs(o + 2, make_float4(v[8], v[9], v[10], v[11]));
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

// grid for a batch of per-cloud segments: y = clouds, x = enough CTAs for the longest segment (threads
// loop only when that would exceed 4 M CTAs)
inline dim3 segment_grid(int64_t chunks_per_cloud, int64_t clouds, int threads) {
  const int64_t gy = std::max<int64_t>(1, std::min<int64_t>(clouds, 65535));
  const int64_t want = std::max<int64_t>(1, ceil_div(chunks_per_cloud, threads));
  // one pass per thread whenever the grid allows it: a cap of a few waves left the T-shape gather with
  // 1.3 loop trips per thread, i.e. a third of the threads idle in the second trip (0.26 of HBM)
  const int64_t cap = std::max<int64_t>(1, (int64_t(1) << 22) / gy);
  return dim3(static_cast<unsigned>(std::min(want, cap)), static_cast<unsigned>(gy));
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
  POPS_CHECK_ARG(reinterpret_cast<uintptr_t>(padded) % 4 == 0 && reinterpret_cast<uintptr_t>(packed) % 4 == 0,
                 "float buffers must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  profile_begin("packed_to_padded", st);
  packed_to_padded_kernel<<<segment_grid(ceil_div(max_size * D + 3, 4), B, 256), 256, 0, st>>>(
      packed, first_idxs, num_inputs, int(B), max_size, int(D), padded);
  profile_end("packed_to_padded", st);
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
  POPS_CHECK_ARG(reinterpret_cast<uintptr_t>(padded) % 4 == 0 && reinterpret_cast<uintptr_t>(packed) % 4 == 0,
                 "float buffers must be 4-byte aligned");
  // a cloud's packed segment is at most num_inputs rows; size x for the padded capacity (the usual case)
  // and let threads loop when first_idxs hands one cloud more rows than that
  profile_begin("padded_to_packed", st);
  padded_to_packed_kernel<<<segment_grid(ceil_div(ceil_div(std::min(max_size, num_inputs) * D + 3, 4) + 1, kCopyUnroll), B, 256), 256, 0, st>>>(
      padded, first_idxs, num_inputs, int(B), max_size, int(D), packed);
  profile_end("padded_to_packed", st);
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
  POPS_CHECK_ARG(L * K * U < (int64_t(1) << 31) - 8 && N < (int64_t(1) << 31), "gather: one cloud's output must stay below 2^31 floats");
  POPS_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 4 == 0,
                 "float buffers must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M == 0) {  // nothing to gather from: every row is zeros (the kernel may always read row 0 of a cloud)
    POPS_CUDA_OK(cudaMemsetAsync(out, 0, size_t(rows) * U * 4, st));
    return POPS_OK;
  }
  const bool v4 = (U % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                  (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const dim3 grid = segment_grid(ceil_div(ceil_div(L * K * U + 3, 4), 4), N, 256);  // 4 chunks per thread and pass
  profile_begin("gather", st);
  if (U == 3 && (L * K) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && reinterpret_cast<uintptr_t>(idx) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(x) % 8 == 0 && get_option("gather_rows3", 1) != 0) {
    // clouds that fit shared memory and are gathered from many times over: stage the cloud once per CTA
    const size_t cloud_bytes = size_t(M) * 12;
    if (cloud_bytes <= 220 * 1024 && (M * 3) % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && L * K >= 4 * M &&
        N <= 65535 && get_option("gather_smem", 1) != 0) {
      const int per_cloud = int(std::max<int64_t>(1, std::min<int64_t>(num_sms() / N, ceil_div(L * K / 4, kGatherSmemThreads))));
      const dim3 gs(static_cast<unsigned>(per_cloud), static_cast<unsigned>(N));
      auto k1 = gather_rows3_smem_kernel<POPS_GATHER_KNN>;
      auto k2 = gather_rows3_smem_kernel<POPS_GATHER_MASKED>;
      auto kern = mode == POPS_GATHER_KNN ? k1 : k2;
      POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cloud_bytes)));
      kern<<<gs, kGatherSmemThreads, cloud_bytes, st>>>(x, idx, lengths, unsigned(L * K), unsigned(K), int(M), int(N), out,
                                                        oob_flag);
      profile_end("gather", st);
      POPS_LAUNCH_OK("gather_rows3_smem_kernel");
      return POPS_OK;
    }
    const dim3 g3 = segment_grid(L * K / 4, N, 256);
    if (mode == POPS_GATHER_KNN)
      gather_rows3_kernel<POPS_GATHER_KNN><<<g3, 256, 0, st>>>(x, idx, lengths, unsigned(L * K), unsigned(K), int(M), int(N), out, oob_flag);
    else
      gather_rows3_kernel<POPS_GATHER_MASKED><<<g3, 256, 0, st>>>(x, idx, lengths, unsigned(L * K), unsigned(K), int(M), int(N), out, oob_flag);
    profile_end("gather", st);
    POPS_LAUNCH_OK("gather_rows3_kernel");
    return POPS_OK;
  }
#define POPS_GATHER(MODE, UT, V4)                                                                        \
  gather_kernel<MODE, UT, V4><<<grid, 256, 0, st>>>(x, idx, lengths, unsigned(L * K), unsigned(K), int(M), \
                                                    unsigned(U), int(N), out, oob_flag)
#define POPS_GATHER_MODE(MODE)                        \
  do {                                                \
    if (v4) POPS_GATHER(MODE, 0, true);               \
    else if (U == 3) POPS_GATHER(MODE, 3, false);     \
    else POPS_GATHER(MODE, 0, false);                 \
  } while (0)
  if (mode == POPS_GATHER_KNN) POPS_GATHER_MODE(POPS_GATHER_KNN); else POPS_GATHER_MODE(POPS_GATHER_MASKED);
#undef POPS_GATHER_MODE
#undef POPS_GATHER
  profile_end("gather", st);
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
