// Spatial ordering pre-pass for the D = 3 KNN scan.
//
// Brute force stays brute force -- every (query, point) pair is still evaluated by the scan --
// but the ORDER in which a warp meets the points decides how often a point beats the running
// K-th distance (a "record"), and records are what the expensive flush path pays for.  With
// points in random order a query sees ~K(1+ln(P/K)) records (~115 buffered groups at P=16384,
// K=16).  If both clouds are sorted along a space-filling curve (Hilbert; Morton kept as option
// knn_curve=0) and every warp starts scanning at its
// own queries' position and moves outward, the first points it meets are already its near
// neighbours and the threshold is tight almost immediately (model: ~20 groups per query).
//
// This file builds that order:
//   1. bbox_maxabs_kernel   per cloud: bounding box of the valid p2 points, max |coord| of p1, p2
//   2. morton_keys_kernel   key = [tensor | cloud | curve code], value = index in cloud (padding
//                           entries: largest code; the stable sort keeps them behind the valid points)
//   3. cub::DeviceRadixSort one sort for every cloud of both tensors
//   4. gather kernels       p2 -> blocks of 64 sorted points (rows x,y,z,w,orig_idx; + sentinels);
//                           p1 -> float4 (x,y,z,orig_idx) in sorted order + each query's home
//                           position in the sorted p2 (binary search of its code)
//   5. box_kernel           bounding box of every block
// Results never depend on the order: the exact 64-bit key (dist, ORIGINAL index) decides.
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>

#include "knn_order.cuh"

namespace pops {

namespace {

inline int clog2(int64_t n) {
  int b = 0;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

struct KeyLayout {
  int axis_bits;     // grid bits per axis
  int code_bits;     // 3 * axis_bits
  int cloud_shift;   // = code_bits (padding entries carry the largest code: see morton_keys_kernel)
  int tensor_shift;  // cloud_shift + clog2(N)
  int end_bit;
};

// Grid resolution follows the cloud size: ~16 cells per point order the blocks of 64 points as
// well as 2^30 cells would, and every 8 key bits less is one radix-sort pass less (each pass is a
// latency-bound ~15 us launch on these small inputs).
inline KeyLayout key_layout(int64_t N, int64_t P2, bool two_tensors) {
  KeyLayout k;
  const int cl = clog2(std::max<int64_t>(N, 1));
  const int want = (clog2(std::max<int64_t>(P2, 1)) + 4 + 2) / 3;
  const int forced = get_option("knn_axis_bits", 0);  // tuning aid
  k.axis_bits = std::min(std::min(10, std::max(1, (30 - cl) / 3)), forced > 0 ? forced : std::max(4, want));
  k.code_bits = 3 * k.axis_bits;
  k.cloud_shift = k.code_bits;
  k.tensor_shift = k.cloud_shift + cl;
  k.end_bit = k.tensor_shift + (two_tensors ? 1 : 0);
  return k;
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : fminf(v, w);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  v = (lane < nw) ? sm[lane] : (is_max ? -FLT_MAX : FLT_MAX);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : fminf(v, w);
  }
  return v;
}

__device__ __forceinline__ unsigned block_reduce_umax(unsigned v, unsigned* sm) {
  v = __reduce_max_sync(0xffffffffu, v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  v = (lane < nw) ? sm[lane] : 0u;
  return __reduce_max_sync(0xffffffffu, v);
}

constexpr int kBboxCluster = 8;   // CTAs per cloud (one thread-block cluster)
constexpr int kBboxThreads = 256;

__device__ __forceinline__ uint32_t bb_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bb_cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split form: arrive at kernel entry, wait right before the first remote store -- a CTA's shared
// memory may only be written from its peers once it is known to have started
__device__ __forceinline__ void bb_cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void bb_cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bb_st_remote(void* local_smem_ptr, uint32_t rank, float value) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(value) : "memory");
}

// One cluster of kBboxCluster CTAs per cloud: every CTA reduces a strided share of the points, the
// partial boxes meet in the shared memory of CTA 0 (DSMEM stores + one cluster barrier).
__global__ void __launch_bounds__(kBboxThreads)
bbox_maxabs_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                   const int64_t* __restrict__ len1, const int64_t* __restrict__ len2, int P1, int P2,
                   bool self_knn, bool union_box, float* __restrict__ bbox,
                   unsigned* __restrict__ maxabs_bits) {
  __shared__ float sm[32];
  __shared__ float part[kBboxCluster][8];  // per CTA: min xyz, max xyz, max |p1|
  bb_cluster_arrive();
  const int n = blockIdx.y;
  const int rank = static_cast<int>(bb_cluster_rank());
  const int t0 = rank * kBboxThreads + threadIdx.x, stride = kBboxCluster * kBboxThreads;
  int64_t L2l = len2[n];
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > P2 ? P2 : L2l));
  const float* b = p2 + static_cast<size_t>(n) * P2 * 3;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned mb = 0u;  // max |coordinate| of everything read, as a bit pattern (+inf and NaN rank highest)
  for (int j = t0; j < L2; j += stride) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = b[static_cast<size_t>(j) * 3 + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
      mb = max(mb, abs_bits(v));
    }
  }
  if (union_box) {  // pair pre-pass: one grid over both clouds
    int64_t L1l = len1[n];
    const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
    const float* a = p1 + static_cast<size_t>(n) * P1 * 3;
    for (int j = t0; j < L1; j += stride) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float v = a[static_cast<size_t>(j) * 3 + d];
        mn[d] = fminf(mn[d], v);
        mx[d] = fmaxf(mx[d], v);
        mb = max(mb, abs_bits(v));
      }
    }
  }
  if (!self_knn && !union_box) {  // p1 is not part of the box, but its magnitude counts
    int64_t L1l = len1[n];
    const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
    const float* a = p1 + static_cast<size_t>(n) * P1 * 3;
    for (int e = t0; e < L1 * 3; e += stride) mb = max(mb, abs_bits(a[e]));
  }
  float out[7];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    out[d] = block_reduce(mn[d], false, sm);
    out[3 + d] = block_reduce(mx[d], true, sm);
  }
  out[6] = __uint_as_float(block_reduce_umax(mb, reinterpret_cast<unsigned*>(sm)));  // bits carried through, never used as a float
  bb_cluster_wait();  // every CTA of the cluster is running: its shared memory can be written
  if (threadIdx.x < 7) {
    float v = out[0];
#pragma unroll
    for (int d = 1; d < 7; ++d) v = (threadIdx.x == d) ? out[d] : v;
    bb_st_remote(&part[rank][threadIdx.x], 0, v);
  }
  bb_cluster_barrier();
  if (rank == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < 7; ++d) out[d] = part[0][d];
    for (int r = 1; r < kBboxCluster; ++r) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        out[d] = fminf(out[d], part[r][d]);
        out[3 + d] = fmaxf(out[3 + d], part[r][3 + d]);
      }
      out[6] = __uint_as_float(max(__float_as_uint(out[6]), __float_as_uint(part[r][6])));
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) bbox[n * 6 + d] = out[d];
    maxabs_bits[n] = __float_as_uint(out[6]);
  }
}

int launch_bbox(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1, int P2,
                bool self_knn, bool union_box, float* bbox, unsigned* maxabs_bits, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kBboxCluster, static_cast<unsigned>(N));
  cfg.blockDim = dim3(kBboxThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kBboxCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, bbox_maxabs_kernel, p1, p2, len1, len2, P1, P2, self_knn, union_box, bbox,
                                  maxabs_bits));
  POPS_LAUNCH_OK("bbox_maxabs_kernel");
  return POPS_OK;
}

// Position of the point's grid cell on a space-filling curve (axis_bits bits per axis).  Hilbert
// (Skilling's transpose form: "Programming the Hilbert curve", AIP Conf. Proc. 707, 2004) rather
// than Morton: consecutive cells of a Hilbert curve are always neighbours, so a run of 64 sorted
// points (a block) or of 128 sorted queries (a warp) has a tighter bounding box and the pruned
// searches visit fewer blocks; the results never depend on the order.
__device__ __forceinline__ unsigned curve_code(const float* p, const float* bb, int axis_bits, bool hilbert) {
  const float cells = static_cast<float>(1u << axis_bits);
  unsigned X[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float lo = bb[d], hi = bb[3 + d];
    const float ext = hi - lo;
    float t = ext > 0.0f ? (p[d] - lo) / ext * cells : 0.0f;
    t = fminf(fmaxf(t, 0.0f), cells - 1.0f);  // clamps p1 points outside p2's box; NaN -> 0
    X[d] = static_cast<unsigned>(t);
  }
  if (!hilbert) return spread3(X[0]) | (spread3(X[1]) << 1) | (spread3(X[2]) << 2);
  const unsigned M = 1u << (axis_bits - 1);
  for (unsigned Q = M; Q > 1u; Q >>= 1) {  // inverse undo
    const unsigned P = Q - 1u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const unsigned t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  X[1] ^= X[0];  // Gray encode
  X[2] ^= X[1];
  unsigned t = 0u;
  for (unsigned Q = M; Q > 1u; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1u;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// Same curve, grid cell from a precomputed scale = cells / extent per axis (0 for a flat axis): a
// multiplication instead of a division per coordinate.  Cells may differ from curve_code's by one at
// cell borders; the order only steers the search, never its results.
__device__ __forceinline__ unsigned curve_code_scaled(float x, float y, float z, const float* lo, const float* scale,
                                                      int axis_bits, bool hilbert) {
  const float top = static_cast<float>((1u << axis_bits) - 1u);
  const float p[3] = {x, y, z};
  unsigned X[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float t = (p[d] - lo[d]) * scale[d];
    t = fminf(fmaxf(t, 0.0f), top);  // NaN -> 0
    X[d] = static_cast<unsigned>(t);
  }
  if (!hilbert) return spread3(X[0]) | (spread3(X[1]) << 1) | (spread3(X[2]) << 2);
  const unsigned M = 1u << (axis_bits - 1);
  for (unsigned Q = M; Q > 1u; Q >>= 1) {  // inverse undo
    const unsigned P = Q - 1u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const unsigned t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  X[1] ^= X[0];  // Gray encode
  X[2] ^= X[1];
  unsigned t = 0u;
  for (unsigned Q = M; Q > 1u; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1u;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// element e in [0, N*P2) -> tensor 0 (p2); [N*P2, N*(P1+P2)) -> tensor 1 (p1)
__global__ void morton_keys_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                   const int64_t* __restrict__ len1, const int64_t* __restrict__ len2,
                                   int N, int P1, int P2, bool self_knn, const float* __restrict__ bbox,
                                   KeyLayout kl, bool hilbert, unsigned* __restrict__ keys,
                                   unsigned* __restrict__ vals) {
  const int64_t total = static_cast<int64_t>(N) * P2 + (self_knn ? 0 : static_cast<int64_t>(N) * P1);
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool second = e >= static_cast<int64_t>(N) * P2;
    const int64_t r = second ? e - static_cast<int64_t>(N) * P2 : e;
    const int P = second ? P1 : P2;
    const int n = static_cast<int>(r / P), j = static_cast<int>(r % P);
    const int64_t L = second ? len1[n] : len2[n];
    const float* src = (second ? p1 : p2) + (static_cast<size_t>(n) * P + j) * 3;
    // padding entries carry the largest code and still end up after every valid point of the cloud:
    // the radix sort is stable and they follow the valid points in the input (j >= L)
    unsigned low = (1u << kl.code_bits) - 1u;
    if (j < L) low = curve_code(src, bbox + n * 6, kl.axis_bits, hilbert);
    keys[e] = (second ? (1u << kl.tensor_shift) : 0u) | (static_cast<unsigned>(n) << kl.cloud_shift) | low;
    vals[e] = static_cast<unsigned>(j);
  }
}

// Sorted p2 in BLOCKS of kBoxPoints points (layout: knn_order.cuh): rows x, y, z, w = |p|^2, the boxes of the
// block's runs (written by box_kernel), the row of original indices -- one contiguous 1408-byte piece per
// block, of which a single TMA bulk copy fetches the 1152 bytes a scan reads.
// Padding entries (beyond lengths2, and the tail of the last block): x = y = z = 0, w = +inf,
// index kNoPoint.
__global__ void gather_p2_kernel(const float* __restrict__ p2, const int64_t* __restrict__ len2, int P2,
                                 int nbox, const unsigned* __restrict__ vals_sorted, bool self_knn,
                                 float* __restrict__ blocks, float4* __restrict__ qsorted,
                                 unsigned* __restrict__ qhome) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nbox * kBoxPoints) return;
  int64_t Ll = len2[n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P2 ? P2 : Ll));
  float x = 0.f, y = 0.f, z = 0.f, w = __int_as_float(0x7f800000);
  unsigned orig = kNoPoint;
  if (j < P2) {
    const unsigned o = vals_sorted[static_cast<size_t>(n) * P2 + j];
    if (j < L) {
      const float* src = p2 + (static_cast<size_t>(n) * P2 + o) * 3;
      x = src[0]; y = src[1]; z = src[2];
      w = fmaf(z, z, fmaf(y, y, x * x));
      orig = o;
    }
    if (self_knn) {
      qsorted[static_cast<size_t>(n) * P2 + j] = make_float4(x, y, z, __uint_as_float(o));
      qhome[static_cast<size_t>(n) * P2 + j] = static_cast<unsigned>(j);
    }
  }
  float* dst = blocks + (static_cast<size_t>(n) * nbox + j / kBoxPoints) * kBlockFloats + (j % kBoxPoints);
  dst[0] = x;
  dst[kBoxPoints] = y;
  dst[2 * kBoxPoints] = z;
  dst[3 * kBoxPoints] = w;
  dst[kIdxOff] = __uint_as_float(orig);
}

__global__ void gather_p1_kernel(const float* __restrict__ p1, const int64_t* __restrict__ len1,
                                 const int64_t* __restrict__ len2, int N, int P1, int P2,
                                 const unsigned* __restrict__ keys_sorted,
                                 const unsigned* __restrict__ vals_sorted, KeyLayout kl,
                                 float4* __restrict__ qsorted, unsigned* __restrict__ qhome) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P1) return;
  int64_t L1l = len1[n], L2l = len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > P1 ? P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > P2 ? P2 : L2l));
  const size_t base1 = static_cast<size_t>(N) * P2 + static_cast<size_t>(n) * P1;
  const unsigned o = vals_sorted[base1 + j];
  float x = 0.f, y = 0.f, z = 0.f;
  unsigned home = 0;
  if (j < L1) {
    const float* src = p1 + (static_cast<size_t>(n) * P1 + o) * 3;
    x = src[0]; y = src[1]; z = src[2];
    // lower bound of this query's (cloud, code) among the sorted keys of p2's cloud n
    const unsigned target = keys_sorted[base1 + j] & ~(1u << kl.tensor_shift);
    const unsigned* k2 = keys_sorted + static_cast<size_t>(n) * P2;
    int lo = 0, hi = L2;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (k2[mid] < target) lo = mid + 1; else hi = mid;
    }
    home = static_cast<unsigned>(lo);
  }
  qsorted[static_cast<size_t>(n) * P1 + j] = make_float4(x, y, z, __uint_as_float(o));
  qhome[static_cast<size_t>(n) * P1 + j] = home;
}

// Lanes hold the 32 points of half `h` of a block (lane = position in the half).  Stores the boxes of
// the half's two runs of kSubPoints points into the block's sub-box area and folds the half into the
// caller's running block box (bmn, bmx: per lane the box of its half-warp; the caller finishes with one
// xor-16 exchange).  Runs without a valid point come out as (+inf, -inf).
__device__ __forceinline__ void half_block_boxes(float x, float y, float z, bool valid, float* block_base, int h,
                                                 int lane, bool store, float (&bmn)[3], float (&bmx)[3]) {
  static_assert(kSubPoints == 16 && kBoxPoints == 64, "two sub-boxes per 32-lane half");
  const float INF = __int_as_float(0x7f800000);
  float mn[3] = {valid ? x : INF, valid ? y : INF, valid ? z : INF};
  float mx[3] = {valid ? x : -INF, valid ? y : -INF, valid ? z : -INF};
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if (store && (lane & 15) == 0) {
    float4* dst = reinterpret_cast<float4*>(block_base + kSubOff) + (h * 2 + (lane >> 4)) * 2;
    dst[0] = make_float4(mn[0], mn[1], mn[2], 0.f);
    dst[1] = make_float4(mx[0], mx[1], mx[2], 0.f);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bmn[d] = fminf(bmn[d], mn[d]);
    bmx[d] = fmaxf(bmx[d], mx[d]);
  }
}

// One warp per block: min / max corner of its valid points (and of its runs of kSubPoints points, stored
// behind the block's rows).  Blocks with no valid point come out as (+inf, -inf) and are never intersected.
__global__ void box_kernel(float* __restrict__ blocks, int nbox, float4* __restrict__ boxes) {
  const int n = blockIdx.y;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= nbox) return;
  const int lane = threadIdx.x & 31;
  float* base = blocks + (static_cast<size_t>(n) * nbox + b) * kBlockFloats;
  const float INF = __int_as_float(0x7f800000);
  float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
#pragma unroll
  for (int h = 0; h < kBoxPoints / 32; ++h) {
    const int i = h * 32 + lane;
    const bool valid = __float_as_uint(base[kIdxOff + i]) != kNoPoint;
    half_block_boxes(base[i], base[kBoxPoints + i], base[2 * kBoxPoints + i], valid, base, h, lane, true, mn, mx);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], 16));
    mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], 16));
  }
  if (lane == 0) {
    float4* dst = boxes + (static_cast<size_t>(n) * nbox + b) * 2;
    dst[0] = make_float4(mn[0], mn[1], mn[2], 0.f);
    dst[1] = make_float4(mx[0], mx[1], mx[2], 0.f);
  }
}

// Pair pre-pass: one sorted cloud serves as the BLOCKS of one direction and as the QUERIES of the
// other.  Entry j of cloud n of the tensor `pts` (sorted position j): block row entry, query entry,
// and the query's home = lower bound of its code among the other tensor's sorted keys.
__global__ void gather_pair_kernel(const float* __restrict__ pts, const int64_t* __restrict__ len_self,
                                   const int64_t* __restrict__ len_other, int P, int P_other, int nbox,
                                   const unsigned* __restrict__ keys_self, const unsigned* __restrict__ vals_self,
                                   const unsigned* __restrict__ keys_other, unsigned self_bit, unsigned other_bit,
                                   float* __restrict__ blocks, float4* __restrict__ qsorted,
                                   unsigned* __restrict__ qhome) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nbox * kBoxPoints) return;
  int64_t Ll = len_self[n], Lol = len_other[n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P ? P : Ll));
  const int Lo = static_cast<int>(Lol < 0 ? 0 : (Lol > P_other ? P_other : Lol));
  float x = 0.f, y = 0.f, z = 0.f, w = __int_as_float(0x7f800000);
  unsigned orig = kNoPoint;
  if (j < P) {
    const unsigned o = vals_self[static_cast<size_t>(n) * P + j];
    unsigned home = 0;
    if (j < L) {
      const float* src = pts + (static_cast<size_t>(n) * P + o) * 3;
      x = src[0]; y = src[1]; z = src[2];
      w = fmaf(z, z, fmaf(y, y, x * x));
      orig = o;
      const unsigned target = (keys_self[static_cast<size_t>(n) * P + j] & ~self_bit) | other_bit;
      const unsigned* ko = keys_other + static_cast<size_t>(n) * P_other;
      int lo = 0, hi = Lo;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ko[mid] < target) lo = mid + 1; else hi = mid;
      }
      home = static_cast<unsigned>(lo);
    }
    qsorted[static_cast<size_t>(n) * P + j] = make_float4(x, y, z, __uint_as_float(o));
    qhome[static_cast<size_t>(n) * P + j] = home;
  }
  float* dst = blocks + (static_cast<size_t>(n) * nbox + j / kBoxPoints) * kBlockFloats + (j % kBoxPoints);
  dst[0] = x;
  dst[kBoxPoints] = y;
  dst[2 * kBoxPoints] = z;
  dst[3 * kBoxPoints] = w;
  dst[kIdxOff] = __uint_as_float(orig);
}

// ---------------------------------------------------------------------------------------------
// Cluster pre-pass: ONE launch, one thread-block cluster per cloud (pair), C CTAs per tensor.
//   grid (T*C, N), cluster (T*C, 1, 1); CTA x of the cluster: tensor t = x / C (0: p2, 1: p1), slice
//   r = x % C of S = 1024 * ITEMS input positions.  Every exchange between the CTAs goes through
//   distributed shared memory; global memory sees the points once on the way in and the ordered
//   buffers once on the way out.
//   1. box + max |coordinate| of the slice, all-to-all over DSMEM (the grid spans both tensors);
//   2. curve codes in registers;
//   3. stable LSD radix sort of (code, index) ACROSS the C CTAs of the tensor, one or two passes of
//      <= 9 bits: per-warp digit counters (match.any ranks inside the warp), prefix over warps, the CTA
//      totals to every peer, bases from the cluster-wide digit histogram, then each key is stored
//      straight into the shared memory of the CTA that owns its destination;
//   4. outputs from the sorted slice: block rows, one bounding box per 64 points (a warp owns whole
//      blocks), sorted queries;
//   5. homes: one lower bound per 16 sorted queries among the OTHER tensor's sorted codes, read from
//      its CTAs' shared memory (the search uses one home per warp, as the start of its outward walk).
// Same order as the multi-launch path (stable by index among equal codes), so the search does the
// same work; 16384-point clouds of the T shape use 4 CTAs each = 128 CTAs on 148 SMs.
// ---------------------------------------------------------------------------------------------
constexpr int kOcThreads = 1024;
constexpr int kOcWarps = kOcThreads / 32;
constexpr int kOcDigitBits = 9;             // 18 code bits in two passes
constexpr int kOcBins = 1 << kOcDigitBits;
constexpr int kOcWcStride = kOcBins + 2;    // u16 row stride of the per-warp counters: column reads hit 32 banks
constexpr int kOcMaxCtas = 8;               // portable cluster size

struct ClusterOrderParams {
  const float* p[2];         // [0] = p2, [1] = p1
  const int64_t* len[2];
  int P[2];
  int mode;                  // 0 self (one tensor), 1 single search (a), 2 pair (a and b)
  int axis_bits, hilbert;
  int C;                     // CTAs per tensor
  int idx_bits;              // log2(C * S): a key is (code << idx_bits) | index
  KnnOrderBuffers a, b;
};

// KeyT: unsigned while code and index fit 32 bits together (18 + 14: clouds of up to 16384 points), else 64 bits
template <int ITEMS, typename KeyT>
struct OcSmem {
  static constexpr int S = kOcThreads * ITEMS;
  KeyT buf[2][S];                            // ping / pong of the radix passes; the idle one stages the local order
  unsigned short wc[kOcWarps][kOcWcStride];  // per-warp digit counts, then their exclusive prefix over warps
  unsigned short ctot[kOcMaxCtas][kOcBins];  // digit totals of every CTA of this tensor (written by the peers)
  unsigned base[kOcBins];                    // first destination of (digit, this CTA) in the tensor's order
  unsigned lbase[kOcBins];                   // first position of the digit in this CTA's local order
  uint2 wsum[kOcWarps];
  float redw[kOcWarps][8];
  float part[kOcMaxCtas][8];                 // per CTA of the cluster: min xyz, max xyz, max |coordinate| bits
  float bb[6];
  float scale[3];                            // grid cells per unit length, per axis
};

__device__ __forceinline__ uint32_t oc_remote(const void* local_smem_ptr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void oc_st(uint32_t addr, unsigned v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void oc_st(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void oc_st_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void oc_ld(uint32_t addr, unsigned& v) {
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
}
__device__ __forceinline__ void oc_ld(uint32_t addr, unsigned long long& v) {
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
}

template <int ITEMS, typename KeyT>
__global__ void __launch_bounds__(kOcThreads, 1)
order_cluster_kernel(const ClusterOrderParams prm) {
  using SM = OcSmem<ITEMS, KeyT>;
  constexpr int S = SM::S;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char ocsm[];
  SM& sm = *reinterpret_cast<SM*>(ocsm);
  bb_cluster_arrive();
  const int C = prm.C;
  const int CT = gridDim.x;  // CTAs per cluster
  const int t = blockIdx.x / C, r = blockIdx.x % C, n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = prm.P[t];
  int64_t Ll = prm.len[t][n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P ? P : Ll));
  const float* pts = prm.p[t] + static_cast<size_t>(n) * P * 3;
  const float INF = __int_as_float(0x7f800000);
  const int IB = prm.idx_bits;
  // local position q of (warp, item i, lane): runs of 32 consecutive points per warp and item; the sort is
  // stable with respect to this order, which is the index order
  const int q0 = warp * (32 * ITEMS) + lane;

  // ---- 1. the slice in registers, box of every tensor of the cluster --------------------------------
  float px[ITEMS], py[ITEMS], pz[ITEMS];
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    unsigned mb = 0u;  // max |coordinate| as a bit pattern (+inf and NaN rank highest)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int j = r * S + q0 + i * 32;
      px[i] = py[i] = pz[i] = 0.0f;
      if (j < L) {
        const float* src = pts + static_cast<size_t>(j) * 3;
        px[i] = src[0]; py[i] = src[1]; pz[i] = src[2];
      }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int j = r * S + q0 + i * 32;
      if (j < L) {
        mn[0] = fminf(mn[0], px[i]); mn[1] = fminf(mn[1], py[i]); mn[2] = fminf(mn[2], pz[i]);
        mx[0] = fmaxf(mx[0], px[i]); mx[1] = fmaxf(mx[1], py[i]); mx[2] = fmaxf(mx[2], pz[i]);
        mb = max(mb, max(abs_bits(px[i]), max(abs_bits(py[i]), abs_bits(pz[i]))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mn[d] = fminf(mn[d], __shfl_xor_sync(FULL, mn[d], o));
        mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL, mx[d], o));
      }
    }
    mb = __reduce_max_sync(FULL, mb);
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        sm.redw[warp][d] = mn[d];
        sm.redw[warp][3 + d] = mx[d];
      }
      sm.redw[warp][6] = __uint_as_float(mb);
    }
    __syncthreads();
    bb_cluster_wait();  // every CTA of the cluster is running: its shared memory can be written
    if (warp < 7) {     // warp k reduces value k over the 32 warps, then lanes < CT send it to every CTA
      float v = sm.redw[lane][warp];
      if (warp < 3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
      } else if (warp < 6) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
      } else {
        v = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(v)));  // bits, never used as a float
      }
      if (lane < CT) oc_st(oc_remote(&sm.part[blockIdx.x][warp], static_cast<uint32_t>(lane)), __float_as_uint(v));
    }
    bb_cluster_barrier();
    if (tid < 3) {
      float lo = sm.part[0][tid], hi = sm.part[0][3 + tid];
      for (int cr = 1; cr < CT; ++cr) {
        lo = fminf(lo, sm.part[cr][tid]);
        hi = fmaxf(hi, sm.part[cr][3 + tid]);
      }
      sm.bb[tid] = lo;
      sm.bb[3 + tid] = hi;
      const float ext = hi - lo;
      sm.scale[tid] = ext > 0.0f ? static_cast<float>(1u << prm.axis_bits) / ext : 0.0f;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
      unsigned m = 0u;
      for (int cr = 0; cr < CT; ++cr) m = max(m, __float_as_uint(sm.part[cr][6]));
#pragma unroll
      for (int d = 0; d < 6; ++d) prm.a.bbox[n * 6 + d] = sm.bb[d];
      prm.a.maxabs_bits[n] = m;
    }
  }

  // ---- 2. keys = (curve code, index) ------------------------------------------------------------------
  const int code_bits = 3 * prm.axis_bits;
  KeyT key[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int j = r * S + q0 + i * 32;
    // padding: largest code, kept behind the valid points by stability (they follow them in the input)
    unsigned code = curve_code_scaled(px[i], py[i], pz[i], sm.bb, sm.scale, prm.axis_bits, prm.hilbert != 0);
    if (j >= L) code = (1u << code_bits) - 1u;
    key[i] = (static_cast<KeyT>(code) << IB) | static_cast<KeyT>(j);
  }

  // ---- 3. stable LSD radix sort across the C CTAs of this tensor --------------------------------------
  const int npass = code_bits > kOcDigitBits ? 2 : 1;
  const int lo_bits = npass == 2 ? code_bits / 2 : code_bits;
  for (int pass = 0; pass < npass; ++pass) {
    const int shift = IB + (pass == 0 ? 0 : lo_bits);
    const int NB = 1 << (pass == 0 ? lo_bits : code_bits - lo_bits);
    KeyT* stage = sm.buf[(pass + 1) & 1];  // the buffer no peer writes during this pass
    if (pass == 1) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) key[i] = sm.buf[0][q0 + i * 32];
    }
    for (int b = lane; b < NB; b += 32) sm.wc[warp][b] = 0;
    __syncwarp();
    unsigned pre[ITEMS];  // keys of this warp with the same digit that come earlier
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const unsigned d = static_cast<unsigned>(key[i] >> shift) & static_cast<unsigned>(NB - 1);
      // lanes with the same digit: one ballot per digit bit (independent, pipelined) beats match.any here
      unsigned peers = FULL;
#pragma unroll
      for (int bit = 0; bit < kOcDigitBits; ++bit) {
        const unsigned bal = __ballot_sync(FULL, (d >> bit) & 1u);
        peers &= ((d >> bit) & 1u) ? bal : ~bal;
      }
      const int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (lane == leader) {
        old = sm.wc[warp][d];
        sm.wc[warp][d] = static_cast<unsigned short>(old + __popc(peers));
      }
      old = __shfl_sync(FULL, old, leader);
      pre[i] = old + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
    }
    __syncthreads();  // (also: every thread has its pass-1 keys in registers before buf[0] becomes the stage)
    // exclusive prefix over the warps, one digit per step and warp (lane = counting warp); the CTA's
    // digit total goes to every peer of the tensor
    for (int b0 = warp * 4; b0 < NB; b0 += kOcWarps * 4) {  // 4 digits per step: independent shuffle chains
      unsigned v[4], incl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = (b0 + u < NB) ? sm.wc[lane][b0 + u] : 0u;
        incl[u] = v[u];
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned up = __shfl_up_sync(FULL, incl[u], o);
          if (lane >= o) incl[u] += up;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b0 + u >= NB) break;
        sm.wc[lane][b0 + u] = static_cast<unsigned short>(incl[u] - v[u]);
        const unsigned tot = __shfl_sync(FULL, incl[u], 31);
        if (lane < C)
          oc_st_u16(oc_remote(&sm.ctot[r][b0 + u], static_cast<uint32_t>(t * C + lane)), static_cast<unsigned short>(tot));
      }
    }
    bb_cluster_barrier();
    {
      // base of (digit, this CTA) = keys of the tensor with a smaller digit + same digit in earlier CTAs;
      // lbase of the digit = keys of this CTA with a smaller digit
      unsigned col = 0, below = 0, own = 0;
      if (tid < NB) {
        for (int cr = 0; cr < C; ++cr) {
          const unsigned v = sm.ctot[cr][tid];
          col += v;
          if (cr < r) below += v;
          if (cr == r) own = v;
        }
      }
      unsigned ic = col, io = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned uc = __shfl_up_sync(FULL, ic, o), uo = __shfl_up_sync(FULL, io, o);
        if (lane >= o) { ic += uc; io += uo; }
      }
      if (lane == 31) sm.wsum[warp] = make_uint2(ic, io);
      __syncthreads();
      if (warp == 0) {
        const uint2 v = sm.wsum[lane];
        unsigned sc = v.x, so = v.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned uc = __shfl_up_sync(FULL, sc, o), uo = __shfl_up_sync(FULL, so, o);
          if (lane >= o) { sc += uc; so += uo; }
        }
        sm.wsum[lane] = make_uint2(sc - v.x, so - v.y);
      }
      __syncthreads();
      if (tid < NB) {
        const uint2 w = sm.wsum[warp];
        sm.base[tid] = w.x + ic - col + below;
        sm.lbase[tid] = w.y + io - own;
      }
      __syncthreads();
    }
    // local order first (plain shared-memory stores), then consecutive threads send consecutive keys:
    // a run of equal digits lands on consecutive addresses of one peer
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const unsigned d = static_cast<unsigned>(key[i] >> shift) & static_cast<unsigned>(NB - 1);
      stage[sm.lbase[d] + sm.wc[warp][d] + pre[i]] = key[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const unsigned lq = static_cast<unsigned>(i * kOcThreads + tid);
      const KeyT k = stage[lq];
      const unsigned d = static_cast<unsigned>(k >> shift) & static_cast<unsigned>(NB - 1);
      const unsigned pos = sm.base[d] + (lq - sm.lbase[d]);
      oc_st(oc_remote(&sm.buf[pass][pos % S], static_cast<uint32_t>(t * C) + pos / S), k);
    }
    bb_cluster_barrier();
  }
  const KeyT* fkey = sm.buf[npass - 1];
  const unsigned idx_mask = (1u << IB) - 1u;

  // ---- 4. outputs from the sorted slice: a warp owns whole blocks of 64 sorted points ---------------
  const bool blocks_role = (t == 0) || prm.mode == 2;           // this tensor is scanned as blocks
  const bool query_role = (t == 1) || prm.mode != 1;            // ... and / or asked as queries
  const KnnOrderBuffers& blk = (t == 0) ? prm.a : prm.b;        // search that scans this tensor
  const KnnOrderBuffers& qry = (t == 1 || prm.mode == 0) ? prm.a : prm.b;  // search that asks it
  const int nbox = static_cast<int>((((P + kBoxPoints - 1) / kBoxPoints) + 31) / 32 * 32);
  for (int bl = warp; bl < S / kBoxPoints; bl += kOcWarps) {
    float bmn[3] = {INF, INF, INF}, bmx[3] = {-INF, -INF, -INF};
    const int gb = r * (S / kBoxPoints) + bl;  // block of the cloud
    const bool wr = blocks_role && gb < nbox;
    float* bbase = blk.blocks + (static_cast<size_t>(n) * nbox + gb) * kBlockFloats;
#pragma unroll
    for (int h = 0; h < kBoxPoints / 32; ++h) {
      const int q = bl * kBoxPoints + h * 32 + lane;
      const int s = r * S + q;
      const unsigned o = static_cast<unsigned>(fkey[q]) & idx_mask;
      float x = 0.f, y = 0.f, z = 0.f, w = INF;
      unsigned orig = kNoPoint;
      if (s < L) {
        const float* src = pts + static_cast<size_t>(o) * 3;
        x = src[0]; y = src[1]; z = src[2];
        w = fmaf(z, z, fmaf(y, y, x * x));
        orig = o;
      }
      if (wr) {
        float* dst = bbase + (h * 32 + lane);
        dst[0] = x;
        dst[kBoxPoints] = y;
        dst[2 * kBoxPoints] = z;
        dst[3 * kBoxPoints] = w;
        dst[kIdxOff] = __uint_as_float(orig);
      }
      if (blocks_role) half_block_boxes(x, y, z, s < L, bbase, h, lane, wr, bmn, bmx);
      if (query_role && s < P) {
        qry.qsorted[static_cast<size_t>(n) * P + s] = make_float4(x, y, z, __uint_as_float(o));
        if (prm.mode == 0) qry.qhome[static_cast<size_t>(n) * P + s] = static_cast<unsigned>(s);
      }
    }
    if (wr) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bmn[d] = fminf(bmn[d], __shfl_xor_sync(FULL, bmn[d], 16));
        bmx[d] = fmaxf(bmx[d], __shfl_xor_sync(FULL, bmx[d], 16));
      }
      if (lane == 0) {
        float4* dst = blk.boxes + (static_cast<size_t>(n) * nbox + gb) * 2;
        dst[0] = make_float4(bmn[0], bmn[1], bmn[2], 0.f);
        dst[1] = make_float4(bmx[0], bmx[1], bmx[2], 0.f);
      }
    }
  }
  if (prm.mode == 0) return;  // no remote access after the last cluster barrier

  // ---- 5. homes: one lower bound per 16 sorted queries, among the other tensor's sorted codes -------
  if (query_role) {
    const int ot = 1 - t;
    const int Po = prm.P[ot];
    int64_t Lol = prm.len[ot][n];
    const int Lo = static_cast<int>(Lol < 0 ? 0 : (Lol > Po ? Po : Lol));
    for (int g = tid; g < S / 16; g += kOcThreads) {
      const int q = g * 16;
      const int s = r * S + q;
      if (s >= P) continue;
      unsigned home = 0;
      if (s < L) {
        const unsigned target = static_cast<unsigned>(fkey[q] >> IB);
        int lo = 0, hi = Lo;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          KeyT v;
          oc_ld(oc_remote(&fkey[mid % S], static_cast<uint32_t>(ot * C + mid / S)), v);
          if (static_cast<unsigned>(v >> IB) < target) lo = mid + 1; else hi = mid;
        }
        home = static_cast<unsigned>(lo);
      }
      unsigned* dst = qry.qhome + static_cast<size_t>(n) * P + s;
      const int cnt = min(16, P - s);
      for (int e = 0; e < cnt; ++e) dst[e] = home;
    }
  }
  bb_cluster_barrier();  // a CTA's shared memory stays alive until its peers have finished reading it
}

template <int ITEMS, typename KeyT>
int launch_cluster_order(const ClusterOrderParams& prm, int N, cudaStream_t st) {
  auto kern = order_cluster_kernel<ITEMS, KeyT>;
  const size_t smem = sizeof(OcSmem<ITEMS, KeyT>);
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int CT = (prm.mode == 0 ? 1 : 2) * prm.C;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(CT), static_cast<unsigned>(N));
  cfg.blockDim = dim3(kOcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(CT);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, prm));
  POPS_LAUNCH_OK("order_cluster_kernel");
  return POPS_OK;
}

template <int ITEMS>
int launch_cluster_order_k(const ClusterOrderParams& prm, int N, cudaStream_t st) {
  if (3 * prm.axis_bits + prm.idx_bits <= 32) return launch_cluster_order<ITEMS, unsigned>(prm, N, st);
  return launch_cluster_order<ITEMS, unsigned long long>(prm, N, st);
}

// mode as in ClusterOrderParams; returns POPS_OK, or -1 when the shape does not fit (caller falls back)
int fused_order(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1, int P2,
                  int mode, const KnnOrderBuffers& a, const KnnOrderBuffers& b, cudaStream_t st) {
  if (get_option("knn_fused_prepass", 1) == 0) return -1;  // 0: the multi-launch path (test / tuning aid)
  const int Pmax = mode == 0 ? P2 : std::max(P1, P2);
  const int T = mode == 0 ? 1 : 2;
  const int cmax = kOcMaxCtas / T;
  // slice size: the smallest (most CTAs per cloud) that still fits the grid into one wave of 1-CTA SMs;
  // if none does, 4096 positions per CTA when the cluster can hold the cloud that way
  const int forced = get_option("knn_cluster_items", 0);
  auto ctas_for = [&](int it) {  // CTAs per tensor with `it` items per thread, 0: does not fit a cluster
    if (forced > 0 && it != forced) return 0;
    int c = 1;
    while (c * kOcThreads * it < Pmax) c *= 2;
    return c <= cmax ? c : 0;
  };
  int items = 0, C = 0;
  for (int it : {2, 4, 8}) {
    const int c = ctas_for(it);
    if (c > 0 && int64_t(N) * T * c <= num_sms()) { items = it; C = c; break; }
  }
  if (items == 0) {
    for (int it : {4, 8, 2}) {
      const int c = ctas_for(it);
      if (c > 0) { items = it; C = c; break; }
    }
  }
  if (items == 0) return -1;
  ClusterOrderParams prm;
  prm.p[0] = p2; prm.p[1] = p1; prm.len[0] = len2; prm.len[1] = len1; prm.P[0] = P2; prm.P[1] = P1;
  prm.mode = mode;
  const int want = (clog2(std::max(Pmax, 1)) + 4 + 2) / 3;
  const int fbits = get_option("knn_axis_bits", 0);
  prm.axis_bits = std::min(6, fbits > 0 ? fbits : std::max(4, want));  // <= 18 code bits: two 9-bit passes
  prm.hilbert = get_option("knn_curve", 1) != 0;
  prm.C = C;
  prm.idx_bits = clog2(int64_t(C) * kOcThreads * items);
  prm.a = a; prm.b = b;
  if (items == 2) return launch_cluster_order_k<2>(prm, N, st);
  if (items == 4) return launch_cluster_order_k<4>(prm, N, st);
  return launch_cluster_order_k<8>(prm, N, st);
}

size_t cub_temp_bytes_for(int64_t items) {
  size_t bytes = 0;
  unsigned* nul = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, nul, nul, nul, nul, static_cast<int>(items), 0, 32);
  return bytes;
}

}  // namespace

size_t knn_order_carve(void* ws, int64_t N, int64_t P1, int64_t P2, KnnOrderBuffers* out) {
  const int64_t items = N * (P1 + P2);
  size_t off = 0;
  char* base = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  KnnOrderBuffers b;
  b.maxabs_bits = reinterpret_cast<unsigned*>(take(size_t(N) * 4));
  b.bbox = reinterpret_cast<float*>(take(size_t(N) * 6 * 4));
  b.blocks = reinterpret_cast<float*>(take(size_t(N) * knn_order_num_boxes(P2) * kBlockFloats * 4));
  b.qsorted = reinterpret_cast<float4*>(take(size_t(N) * P1 * 16));
  b.qhome = reinterpret_cast<unsigned*>(take(size_t(N) * P1 * 4));
  b.boxes = reinterpret_cast<float4*>(take(size_t(N) * knn_order_num_boxes(P2) * 32));
  b.keys_in = reinterpret_cast<unsigned*>(take(size_t(items) * 4));
  b.keys_out = reinterpret_cast<unsigned*>(take(size_t(items) * 4));
  b.vals_in = reinterpret_cast<unsigned*>(take(size_t(items) * 4));
  b.vals_out = reinterpret_cast<unsigned*>(take(size_t(items) * 4));
  b.cub_temp_bytes = cub_temp_bytes_for(items);
  b.cub_temp = take(b.cub_temp_bytes);
  if (out) *out = b;
  return off;
}

size_t knn_order_workspace_bytes(int64_t N, int64_t P1, int64_t P2) {
  return knn_order_carve(nullptr, N, P1, P2, nullptr) + 256;
}

int knn_order_prepass(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2,
                      int N, int P1, int P2, bool self_knn, const KnnOrderBuffers& b, cudaStream_t st) {
  struct Prof {  // per-call device time of the whole pre-pass under pops_profile_enable(1)
    cudaStream_t s;
    explicit Prof(cudaStream_t s_) : s(s_) { profile_begin("knn_order", s); }
    ~Prof() { profile_end("knn_order", s); }
  } prof(st);
  {
    const int rc = fused_order(p1, p2, len1, len2, N, P1, P2, self_knn ? 0 : 1, b, b, st);
    if (rc >= 0) return rc;
  }
  const int nbox = static_cast<int>(knn_order_num_boxes(P2));
  const KeyLayout kl = key_layout(N, P2, !self_knn);
  {
    const int rc = launch_bbox(p1, p2, len1, len2, N, P1, P2, self_knn, false, b.bbox, b.maxabs_bits, st);
    if (rc != POPS_OK) return rc;
  }
  const int64_t items = static_cast<int64_t>(N) * P2 + (self_knn ? 0 : static_cast<int64_t>(N) * P1);
  {
    const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(items, 256), int64_t(num_sms()) * 16));
    morton_keys_kernel<<<blocks, 256, 0, st>>>(p1, p2, len1, len2, N, P1, P2, self_knn, b.bbox, kl,
                                               get_option("knn_curve", 1) != 0, b.keys_in, b.vals_in);
    POPS_LAUNCH_OK("morton_keys_kernel");
  }
  size_t temp = b.cub_temp_bytes;
  POPS_CUDA_OK(cub::DeviceRadixSort::SortPairs(b.cub_temp, temp, b.keys_in, b.keys_out, b.vals_in,
                                               b.vals_out, static_cast<int>(items), 0, kl.end_bit, st));
  g_launch_count.fetch_add(4, std::memory_order_relaxed);  // cub: histogram + onesweep passes
  {
    dim3 grid(static_cast<unsigned>(ceil_div(int64_t(nbox) * kBoxPoints, 256)), N);
    gather_p2_kernel<<<grid, 256, 0, st>>>(p2, len2, P2, nbox, b.vals_out, self_knn, b.blocks, b.qsorted,
                                           b.qhome);
    POPS_LAUNCH_OK("gather_p2_kernel");
  }
  {
    dim3 grid(static_cast<unsigned>(ceil_div(nbox, 8)), N);
    box_kernel<<<grid, 256, 0, st>>>(b.blocks, nbox, b.boxes);
    POPS_LAUNCH_OK("box_kernel");
  }
  if (!self_knn) {
    dim3 grid(static_cast<unsigned>(ceil_div(P1, 256)), N);
    gather_p1_kernel<<<grid, 256, 0, st>>>(p1, len1, len2, N, P1, P2, b.keys_out, b.vals_out, kl,
                                           b.qsorted, b.qhome);
    POPS_LAUNCH_OK("gather_p1_kernel");
  }
  return POPS_OK;
}

// Both directions of a two-sided search (chamfer: x -> y and y -> x) from ONE sort: `a` serves the
// queries p1 over the blocks of p2, `b` the queries p2 over the blocks of p1.  Only a's scratch
// buffers are used; b.maxabs_bits / b.bbox are not written (the caller points them at a's).
int knn_order_prepass_pair(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                           int P1, int P2, const KnnOrderBuffers& a, const KnnOrderBuffers& b, cudaStream_t st) {
  struct Prof {
    cudaStream_t s;
    explicit Prof(cudaStream_t s_) : s(s_) { profile_begin("knn_order", s); }
    ~Prof() { profile_end("knn_order", s); }
  } prof(st);
  {
    const int rc = fused_order(p1, p2, len1, len2, N, P1, P2, 2, a, b, st);
    if (rc >= 0) return rc;
  }
  const int nbox2 = static_cast<int>(knn_order_num_boxes(P2)), nbox1 = static_cast<int>(knn_order_num_boxes(P1));
  const KeyLayout kl = key_layout(N, std::max(P1, P2), true);
  {
    const int rc = launch_bbox(p1, p2, len1, len2, N, P1, P2, false, true, a.bbox, a.maxabs_bits, st);
    if (rc != POPS_OK) return rc;
  }
  const int64_t items = static_cast<int64_t>(N) * (P1 + P2);
  {
    const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(items, 256), int64_t(num_sms()) * 16));
    morton_keys_kernel<<<blocks, 256, 0, st>>>(p1, p2, len1, len2, N, P1, P2, false, a.bbox, kl,
                                               get_option("knn_curve", 1) != 0, a.keys_in, a.vals_in);
    POPS_LAUNCH_OK("morton_keys_kernel");
  }
  size_t temp = a.cub_temp_bytes;
  POPS_CUDA_OK(cub::DeviceRadixSort::SortPairs(a.cub_temp, temp, a.keys_in, a.keys_out, a.vals_in, a.vals_out,
                                               static_cast<int>(items), 0, kl.end_bit, st));
  g_launch_count.fetch_add(4, std::memory_order_relaxed);
  const unsigned tbit = 1u << kl.tensor_shift;
  const size_t base1 = static_cast<size_t>(N) * P2;  // sorted p1 entries follow the p2 entries
  {
    dim3 grid(static_cast<unsigned>(ceil_div(int64_t(nbox2) * kBoxPoints, 256)), N);
    gather_pair_kernel<<<grid, 256, 0, st>>>(p2, len2, len1, P2, P1, nbox2, a.keys_out, a.vals_out,
                                             a.keys_out + base1, 0u, tbit, a.blocks, b.qsorted, b.qhome);
    POPS_LAUNCH_OK("gather_pair_kernel");
  }
  {
    dim3 grid(static_cast<unsigned>(ceil_div(int64_t(nbox1) * kBoxPoints, 256)), N);
    gather_pair_kernel<<<grid, 256, 0, st>>>(p1, len1, len2, P1, P2, nbox1, a.keys_out + base1, a.vals_out + base1,
                                             a.keys_out, tbit, 0u, b.blocks, a.qsorted, a.qhome);
    POPS_LAUNCH_OK("gather_pair_kernel");
  }
  {
    dim3 grid(static_cast<unsigned>(ceil_div(nbox2, 8)), N);
    box_kernel<<<grid, 256, 0, st>>>(a.blocks, nbox2, a.boxes);
    POPS_LAUNCH_OK("box_kernel");
  }
  {
    dim3 grid(static_cast<unsigned>(ceil_div(nbox1, 8)), N);
    box_kernel<<<grid, 256, 0, st>>>(b.blocks, nbox1, b.boxes);
    POPS_LAUNCH_OK("box_kernel");
  }
  return POPS_OK;
}

}  // namespace pops
