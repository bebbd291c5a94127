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
#include <cub/block/block_radix_sort.cuh>
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

// Sorted p2 in BLOCKS of kBoxPoints points: [n][block][row][kBoxPoints], rows x, y, z, w = |p|^2,
// original index -- one contiguous 1280-byte piece per block, fetched by a single TMA bulk copy.
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
  dst[4 * kBoxPoints] = __uint_as_float(orig);
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

// One warp per block: min / max corner of its valid points.  Blocks with no valid point come out
// as (+inf, -inf) and are never intersected.
__global__ void box_kernel(const float* __restrict__ blocks, int nbox, float4* __restrict__ boxes) {
  const int n = blockIdx.y;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= nbox) return;
  const int lane = threadIdx.x & 31;
  const float* base = blocks + (static_cast<size_t>(n) * nbox + b) * kBlockFloats;
  const float INF = __int_as_float(0x7f800000);
  float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
  for (int i = lane; i < kBoxPoints; i += 32) {
    if (__float_as_uint(base[4 * kBoxPoints + i]) == kNoPoint) continue;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = base[d * kBoxPoints + i];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
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
  dst[4 * kBoxPoints] = __uint_as_float(orig);
}

// ---------------------------------------------------------------------------------------------
// Fused pre-pass for clouds that fit one CTA (<= 1024 * ITEMS = 8192 points per tensor): ONE launch does
// what bbox + keys + device radix sort + gathers + boxes (11-14 launches) do -- these shapes are
// launch-bound (chamfer: 26 launches per step), and a 16 K-key sort is a shared-memory job.
//   grid (T, N), cluster (T, 1, 1): CTA t of the cluster orders tensor t of cloud n (t = 0: p2, the
//   blocks of search `a`; t = 1: p1, the queries of `a` and, in pair mode, the blocks of `b`).
//   1. box + max |coordinate| of the CTA's own points; the two CTAs exchange them over DSMEM (the
//      grid of the curve codes spans both clouds);
//   2. curve codes, cub::BlockRadixSort on (code, index) in shared memory (stable: padding entries
//      carry the largest code and follow the valid points);
//   3. straight from the sorted registers: block rows, per-block boxes (shuffle reduction over the
//      64 / ITEMS threads of a block), sorted queries, sorted codes;
//   4. cluster barrier, then every query's home = lower bound of its code among the OTHER tensor's
//      sorted codes.
// ---------------------------------------------------------------------------------------------
constexpr int kFusedThreads = 1024;

struct FusedOrderParams {
  const float* p[2];         // [0] = p2, [1] = p1
  const int64_t* len[2];
  int P[2];
  int mode;                  // 0 self (one tensor), 1 single search (a), 2 pair (a and b)
  int axis_bits, hilbert;
  KnnOrderBuffers a, b;
  unsigned* codes[2];        // sorted codes per tensor: [N][P[t]]
};

template <int ITEMS>
__global__ void __launch_bounds__(kFusedThreads, 1)
order_cloud_kernel(const FusedOrderParams prm) {
  // (code << IDX_BITS | index) in ONE 32-bit key, sorted on the code bits only: no value array to carry
  constexpr int IDX_BITS = ITEMS == 8 ? 13 : 14;  // log2(1024 * ITEMS)
  using Sort = cub::BlockRadixSort<unsigned, kFusedThreads, ITEMS>;
  extern __shared__ __align__(16) unsigned char fsm[];
  typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(fsm);
  __shared__ float red[32];
  __shared__ float part[2][8];  // per CTA of the cluster: min xyz, max xyz
  __shared__ float bb[6];
  bb_cluster_arrive();
  const int T = gridDim.x;
  const int t = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const int P = prm.P[t];
  int64_t Ll = prm.len[t][n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P ? P : Ll));
  const float* pts = prm.p[t] + static_cast<size_t>(n) * P * 3;

  // ---- 1. box of both tensors ---------------------------------------------------------------------
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned mb = 0u;  // max |coordinate| as a bit pattern (+inf and NaN rank highest)
  for (int j = tid; j < L; j += kFusedThreads) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pts[static_cast<size_t>(j) * 3 + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
      mb = max(mb, abs_bits(v));
    }
  }
  float out[7];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    out[d] = block_reduce(mn[d], false, red);
    out[3 + d] = block_reduce(mx[d], true, red);
  }
  out[6] = __uint_as_float(block_reduce_umax(mb, reinterpret_cast<unsigned*>(red)));  // bits, never used as a float
  bb_cluster_wait();  // the peer CTA is running: its shared memory can be written
  if (tid < 7) {
    float v = out[0];
#pragma unroll
    for (int d = 1; d < 7; ++d) v = (tid == d) ? out[d] : v;
    for (int r = 0; r < T; ++r) bb_st_remote(&part[t][tid], static_cast<uint32_t>(r), v);
  }
  if (T > 1) bb_cluster_barrier(); else __syncthreads();
  if (tid < 3) {
    float lo = part[0][tid], hi = part[0][3 + tid];
    if (T > 1) {
      lo = fminf(lo, part[1][tid]);
      hi = fmaxf(hi, part[1][3 + tid]);
    }
    bb[tid] = lo;
    bb[3 + tid] = hi;
  }
  __syncthreads();
  if (t == 0 && tid == 0) {
    unsigned m = __float_as_uint(part[0][6]);
    if (T > 1) m = max(m, __float_as_uint(part[1][6]));
#pragma unroll
    for (int d = 0; d < 6; ++d) prm.a.bbox[n * 6 + d] = bb[d];
    prm.a.maxabs_bits[n] = m;
  }

  // ---- 2. codes + sort (blocked arrangement: thread tid holds positions tid*ITEMS + i) ------------
  const int code_bits = 3 * prm.axis_bits;
  unsigned keys[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int j = tid * ITEMS + i;
    unsigned code = (1u << code_bits) - 1u;  // padding: largest code, kept behind the valid points by stability
    if (j < L) code = curve_code(pts + static_cast<size_t>(j) * 3, bb, prm.axis_bits, prm.hilbert != 0);
    keys[i] = (code << IDX_BITS) | static_cast<unsigned>(j);
  }
  Sort(temp).Sort(keys, IDX_BITS, IDX_BITS + code_bits);

  // ---- 3. outputs from the sorted registers -----------------------------------------------------
  const bool blocks_role = (t == 0) || prm.mode == 2;           // this tensor is scanned as blocks
  const bool query_role = (t == 1) || prm.mode != 1;            // ... and / or asked as queries
  const KnnOrderBuffers& blk = (t == 0) ? prm.a : prm.b;        // search that scans this tensor
  const KnnOrderBuffers& qry = (t == 1 || prm.mode == 0) ? prm.a : prm.b;  // search that asks it
  const int nbox = static_cast<int>((((P + kBoxPoints - 1) / kBoxPoints) + 31) / 32 * 32);
  const float INF = __int_as_float(0x7f800000);
  float bmn[3] = {INF, INF, INF}, bmx[3] = {-INF, -INF, -INF};
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int s = tid * ITEMS + i;
    const unsigned o = keys[i] & ((1u << IDX_BITS) - 1u);
    float x = 0.f, y = 0.f, z = 0.f, w = INF;
    unsigned orig = kNoPoint;
    if (s < L) {
      const float* src = pts + static_cast<size_t>(o) * 3;
      x = src[0]; y = src[1]; z = src[2];
      w = fmaf(z, z, fmaf(y, y, x * x));
      orig = o;
      bmn[0] = fminf(bmn[0], x); bmn[1] = fminf(bmn[1], y); bmn[2] = fminf(bmn[2], z);
      bmx[0] = fmaxf(bmx[0], x); bmx[1] = fmaxf(bmx[1], y); bmx[2] = fmaxf(bmx[2], z);
    }
    if (blocks_role && s < nbox * kBoxPoints) {
      float* dst = blk.blocks + (static_cast<size_t>(n) * nbox + s / kBoxPoints) * kBlockFloats + (s % kBoxPoints);
      dst[0] = x;
      dst[kBoxPoints] = y;
      dst[2 * kBoxPoints] = z;
      dst[3 * kBoxPoints] = w;
      dst[4 * kBoxPoints] = __uint_as_float(orig);
    }
    if (s < P) {
      if (query_role) qry.qsorted[static_cast<size_t>(n) * P + s] = make_float4(x, y, z, __uint_as_float(o));
      prm.codes[t][static_cast<size_t>(n) * P + s] = keys[i] >> IDX_BITS;
    }
  }
  if (blocks_role) {
    constexpr int TPB = kBoxPoints / ITEMS;  // threads per block of 64 sorted points (consecutive lanes)
    static_assert(TPB >= 1 && TPB <= 32 && (TPB & (TPB - 1)) == 0, "a block is a power-of-two run of lanes");
#pragma unroll
    for (int o2 = TPB / 2; o2 > 0; o2 >>= 1) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bmn[d] = fminf(bmn[d], __shfl_xor_sync(0xffffffffu, bmn[d], o2));
        bmx[d] = fmaxf(bmx[d], __shfl_xor_sync(0xffffffffu, bmx[d], o2));
      }
    }
    const int b0 = tid / TPB;
    if (tid % TPB == 0 && b0 < nbox) {
      float4* dst = blk.boxes + (static_cast<size_t>(n) * nbox + b0) * 2;
      dst[0] = make_float4(bmn[0], bmn[1], bmn[2], 0.f);
      dst[1] = make_float4(bmx[0], bmx[1], bmx[2], 0.f);
    }
  }

  // ---- 4. homes: lower bound of every query's code among the other tensor's sorted codes ----------
  if (prm.mode == 0) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int s = tid * ITEMS + i;
      if (s < P) prm.a.qhome[static_cast<size_t>(n) * P + s] = static_cast<unsigned>(s);
    }
    return;
  }
  __threadfence();
  bb_cluster_barrier();
  if (!query_role) return;
  const int ot = 1 - t;
  const int Po = prm.P[ot];
  int64_t Lol = prm.len[ot][n];
  const int Lo = static_cast<int>(Lol < 0 ? 0 : (Lol > Po ? Po : Lol));
  const unsigned* ko = prm.codes[ot] + static_cast<size_t>(n) * Po;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int s = tid * ITEMS + i;
    if (s >= P) continue;
    unsigned home = 0;
    if (s < L) {
      int lo = 0, hi = Lo;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(ko + mid) < (keys[i] >> IDX_BITS)) lo = mid + 1; else hi = mid;
      }
      home = static_cast<unsigned>(lo);
    }
    qry.qhome[static_cast<size_t>(n) * P + s] = home;
  }
}

template <int ITEMS>
int launch_fused_order(const FusedOrderParams& prm, int N, cudaStream_t st) {
  using Sort = cub::BlockRadixSort<unsigned, kFusedThreads, ITEMS>;
  auto kern = order_cloud_kernel<ITEMS>;
  const size_t smem = sizeof(typename Sort::TempStorage);
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int T = prm.mode == 0 ? 1 : 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(T, static_cast<unsigned>(N));
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = T;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, prm));
  POPS_LAUNCH_OK("order_cloud_kernel");
  return POPS_OK;
}

// mode as in FusedOrderParams; returns POPS_OK, or -1 when the shape does not fit (caller falls back)
int fused_order(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1, int P2,
                int mode, const KnnOrderBuffers& a, const KnnOrderBuffers& b, cudaStream_t st) {
  const int Pmax = mode == 0 ? P2 : std::max(P1, P2);
  // up to 8 items per thread by default (knn_fused_items = 16 admits clouds of up to 16384 points): N or
  // 2N CTAs leave most of the 148 SMs idle, which costs more GPU time than the launches it saves once the
  // clouds are large (T shape, N = 32: 0.96 vs 0.915 ms per call); up to 8192 points the single launch
  // wins on hosts where the step is launch-bound (chamfer step 0.50 vs 0.60 ms) and costs ~4 % where it
  // is not (0.446 vs 0.427 ms)
  const int items = get_option("knn_fused_items", 8) <= 8 ? 8 : 16;  // the two compiled forms; nothing larger exists
  if (Pmax > kFusedThreads * items || get_option("knn_fused_prepass", 1) == 0) return -1;
  FusedOrderParams prm;
  prm.p[0] = p2; prm.p[1] = p1; prm.len[0] = len2; prm.len[1] = len1; prm.P[0] = P2; prm.P[1] = P1;
  prm.mode = mode;
  const int want = (clog2(std::max(Pmax, 1)) + 4 + 2) / 3;
  const int forced = get_option("knn_axis_bits", 0);
  prm.axis_bits = std::min(6, forced > 0 ? forced : std::max(4, want));  // 18 code bits + 13-14 index bits = one key
  prm.hilbert = get_option("knn_curve", 1) != 0;
  prm.a = a; prm.b = b;
  prm.codes[0] = a.keys_out;                                   // [N][P2]
  prm.codes[1] = a.keys_out + static_cast<size_t>(N) * P2;     // [N][P1]
  if (Pmax <= kFusedThreads * 8) return launch_fused_order<8>(prm, N, st);
  return launch_fused_order<16>(prm, N, st);
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
