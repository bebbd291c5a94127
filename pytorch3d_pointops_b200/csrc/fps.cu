// Farthest point sampling for sm_100a.
//
// Replaces csrc/sample_farthest_points/{.cu,_cpu.cpp}.  Contract = the reference CPU path
// (sample_farthest_points_cpu.cpp:14-103): idx[n,0] = start_idxs[n]; then batch_k-1 times
//   mind[p] = min(mind[p], dist2(last, p))   (unfused f32, (last - p)^2 summed d = 0..D-1)
//   last    = FIRST maximum of mind          (lowest index on ties)
// (forcing selected points to 0 in the reference, :68-72, is the same thing: dist2(p,p) = 0).
//
// D == 3 design: the whole iteration state lives in REGISTERS.  One thread-block CLUSTER of C
// CTAs (C in {1,2,4,8,16}) owns one cloud; each of the C*1024 threads keeps PT points
// (x, y, z, mind) in registers for the entire run, so an iteration touches no global or shared
// memory for point data -- the reference re-reads 20 B/point/iteration from global memory
// (sample_farthest_points.cu:63-76).  The arg-max is: REDUX warp reduce -> 32 warp slots in
// shared memory -> one warp -> the CTA's winner (key, index AND coordinates) is pushed into
// every CTA of the cluster through distributed shared memory -> one cluster barrier -> every
// thread picks the cluster winner locally.  2 barriers per iteration, no global round trip.
//
// Generic D: one CTA per cloud, mind in a global scratch row (L2 resident), same reduction.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace pops {

constexpr int kFpsThreads = 1024;  // generic-D kernel
constexpr int kFpsWarps = kFpsThreads / 32;
constexpr int kFps3Threads = 256;  // D = 3 kernel: few fat threads (<= 32 points each in registers)
constexpr int kFps3Warps = kFps3Threads / 32;
constexpr int kFps3MaxPT = 32;

struct FpsSlot {  // 32 bytes
  int key;        // float bits of mind (>= 0) or negative when the CTA has no valid point
  int idx;
  float x, y, z;
  int pad0, pad1, pad2;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, int a, int b, int c, int d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// 16 bytes into a peer CTA's shared memory, completing on THAT CTA's mbarrier (both addresses shared::cluster)
__device__ __forceinline__ void st_async_v4(uint32_t addr, int a, int b, int c, int d, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar)
               : "memory");
}

// warp arg-max with first-index tie break.  Returns true in the winning lane.
__device__ __forceinline__ bool warp_argmax(int key, int idx, int* wkey, int* widx) {
  const int mk = __reduce_max_sync(0xffffffffu, key);
  const int cand = (key == mk) ? idx : 0x7fffffff;
  const int mi = __reduce_min_sync(0xffffffffu, cand);
  *wkey = mk;
  *widx = mi;
  return key == mk && idx == mi;
}

template <int PT>
__global__ void __launch_bounds__(kFps3Threads, 1)
fps_d3_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths,
              const int64_t* __restrict__ Ks, const int64_t* __restrict__ start_idxs, int P,
              int max_K, int C, int push, int64_t* __restrict__ out) {
  __shared__ __align__(16) FpsSlot wslots[2][kFps3Warps];  // per-warp winners, double buffered
  // winners of the cluster, double buffered: one per CTA (push = 0) or one per warp of every CTA (push = 1)
  __shared__ __align__(16) FpsSlot cslots[2][16 * kFps3Warps];
  __shared__ __align__(8) uint64_t bars[2];  // push = 1: "all slots of this parity have landed"
  // push = 1: the CTA's points again, [3][PT * threads]: the winner's coordinates are read by index instead
  // of being selected out of the owning lane's registers (3 x PT selects per iteration)
  extern __shared__ __align__(16) float spts[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (C > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int n = blockIdx.x / C;
  int64_t Ll = lengths[n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P ? P : Ll));
  int64_t kl = Ks[n];
  const int kn = static_cast<int>(kl < L ? (kl < 0 ? 0 : kl) : L);  // min(lengths, K)
  const float* pts = points + static_cast<size_t>(n) * P * 3;
  int64_t* o = out + static_cast<size_t>(n) * max_K;

  int last = static_cast<int>(start_idxs[n]);
  if (rank == 0) {
    for (int k = tid; k < max_K; k += kFps3Threads)
      if (k == 0) o[0] = last; else if (k >= kn) o[k] = -1;
  }
  if (kn <= 1) return;  // uniform across the cluster: nothing else to select
  last = min(max(last, 0), L - 1);

  // point i of this thread has index  i*stride + base  (indices ascend with i)
  const int stride = C * kFps3Threads;  // a power of two
  const int lstride = 31 - __clz(stride);
  const int base = rank * kFps3Threads + tid;
  // coordinates are kept NEGATED: last - p = last + (-p) exactly, and the packed add takes no negation
  float x[PT], y[PT], z[PT], mind[PT];
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    const int p = i * stride + base;
    const bool v = p < L;
    const float vx = v ? pts[static_cast<size_t>(p) * 3 + 0] : 0.0f;
    const float vy = v ? pts[static_cast<size_t>(p) * 3 + 1] : 0.0f;
    const float vz = v ? pts[static_cast<size_t>(p) * 3 + 2] : 0.0f;
    x[i] = -vx; y[i] = -vy; z[i] = -vz;
    mind[i] = v ? FLT_MAX : -1.0f;  // -1: never the maximum, min() keeps it
    if (C > 1 && push) {
      spts[i * kFps3Threads + tid] = vx;
      spts[(PT + i) * kFps3Threads + tid] = vy;
      spts[(2 * PT + i) * kFps3Threads + tid] = vz;
    }
  }
  float lx = pts[static_cast<size_t>(last) * 3 + 0];
  float ly = pts[static_cast<size_t>(last) * 3 + 1];
  float lz = pts[static_cast<size_t>(last) * 3 + 2];

  const uint32_t slot_bytes = static_cast<uint32_t>(C) * kFps3Warps * 32u;  // what one iteration delivers to a CTA
  if (C > 1 && push) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(&bars[1], slot_bytes);             // iteration 1
      if (kn > 2) mbar_arrive_expect_tx(&bars[0], slot_bytes);  // iteration 2
    }
  }
  if (C > 1) cluster_barrier();  // every CTA of the cluster is running before the first remote store
  for (int k = 1; k < kn; ++k) {
    const int par = k & 1;
    // ---- update min-distances; track only the maximum VALUE in the dense loop -----------------
    float best = -1.0f;
    if (PT >= 2) {
#pragma unroll
      for (int i = 0; i < PT; i += 2) {
        const int i1 = (i + 1 < PT) ? i + 1 : i;
        // packed subtract / multiply (IEEE, unfused), scalar adds (ptxas would fuse mul2+add2)
        const float2 dx = __fadd2_rn(make_float2(lx, lx), make_float2(x[i], x[i1]));
        const float2 dy = __fadd2_rn(make_float2(ly, ly), make_float2(y[i], y[i1]));
        const float2 dz = __fadd2_rn(make_float2(lz, lz), make_float2(z[i], z[i1]));
        const float2 xx = __fmul2_rn(dx, dx), yy = __fmul2_rn(dy, dy), zz = __fmul2_rn(dz, dz);
        const float d0 = __fadd_rn(__fadd_rn(xx.x, yy.x), zz.x);
        const float d1 = __fadd_rn(__fadd_rn(xx.y, yy.y), zz.y);
        mind[i] = fminf(mind[i], d0);
        mind[i1] = fminf(mind[i1], d1);
        best = fmaxf(best, fmaxf(mind[i], mind[i1]));
      }
    } else {
      const float dx = __fadd_rn(lx, x[0]), dy = __fadd_rn(ly, y[0]), dz = __fadd_rn(lz, z[0]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      mind[0] = fminf(mind[0], d);
      best = mind[0];
    }
    // ---- warp arg-max: value first, the index only for lanes that hold the maximum ------------
    const int bkey = __float_as_int(best);
    const int wkey = __reduce_max_sync(0xffffffffu, bkey);
    int cand = 0x7fffffff, ci = 0;
    if (bkey == wkey) {
#pragma unroll
      for (int i = PT - 1; i >= 0; --i)
        if (mind[i] == best) ci = i;  // lowest local slot holding the maximum
      cand = ci * stride + base;
    }
    const int widx = __reduce_min_sync(0xffffffffu, cand);
    if (C > 1 && push) {
      // One-sided exchange, no CTA or cluster barrier: every WARP sends its winner to every CTA of the
      // cluster (lanes 0..15: first half of the slot to CTA `lane`, lanes 16..31: second half), each
      // store completing 16 bytes on the receiver's mbarrier; a CTA waits on its own mbarrier only.
      // Slot reuse is safe without further synchronisation: a peer sends iteration k+2 only after it
      // has received iteration k+1 from every warp here, and a warp sends k+1 after it has read k.
      // the warp's winner is one of this CTA's points: index -> (slot i, thread) -> shared-memory copy
      const int wi = widx >> lstride, wt = (widx & (stride - 1)) - rank * kFps3Threads;
      const float bx = spts[wi * kFps3Threads + wt];
      const float by = spts[(PT + wi) * kFps3Threads + wt];
      const float bz = spts[(2 * PT + wi) * kFps3Threads + wt];
      const int dst = lane & 15;
      if (dst < C) {
        const uint32_t rs = map_to_cta(smem_u32(&cslots[par][rank * kFps3Warps + warp]), static_cast<uint32_t>(dst));
        const uint32_t rb = map_to_cta(smem_u32(&bars[par]), static_cast<uint32_t>(dst));
        if (lane < 16) st_async_v4(rs, wkey, widx, __float_as_int(bx), __float_as_int(by), rb);
        else st_async_v4(rs + 16, __float_as_int(bz), 0, 0, 0, rb);
      }
      mbar_wait(&bars[par], static_cast<uint32_t>((k - 1) >> 1) & 1u);
      if (tid == 0 && k + 2 < kn) mbar_arrive_expect_tx(&bars[par], slot_bytes);  // arm this parity for iteration k + 2
      const int nslots = C * kFps3Warps;
      int bk = static_cast<int>(0x80000000u), bi = 0x7fffffff, bs = 0;
#pragma unroll
      for (int u = 0; u < (16 * kFps3Warps) / 32; ++u) {
        const int sl = lane + u * 32;
        if (sl < nslots) {
          const int2 ki = *reinterpret_cast<const int2*>(&cslots[par][sl]);
          if (ki.x > bk || (ki.x == bk && ki.y < bi)) { bk = ki.x; bi = ki.y; bs = sl; }
        }
      }
      const int ck = __reduce_max_sync(0xffffffffu, bk);
      const int cidx = __reduce_min_sync(0xffffffffu, bk == ck ? bi : 0x7fffffff);
      const int wl = __ffs(__ballot_sync(0xffffffffu, bk == ck && bi == cidx)) - 1;
      const int ws = __shfl_sync(0xffffffffu, bs, wl);
      lx = cslots[par][ws].x; ly = cslots[par][ws].y; lz = cslots[par][ws].z;
      if (rank == 0 && tid == 0) o[k] = cidx;
      continue;
    }
    if (cand == widx) {  // exactly one lane (indices are unique)
      float bx = x[0], by = y[0], bz = z[0];
#pragma unroll
      for (int i = 1; i < PT; ++i)
        if (ci == i) { bx = x[i]; by = y[i]; bz = z[i]; }
      FpsSlot s;
      s.key = wkey; s.idx = widx; s.x = -bx; s.y = -by; s.z = -bz; s.pad0 = s.pad1 = s.pad2 = 0;
      wslots[par][warp] = s;
    }
    __syncthreads();

    FpsSlot b;
    if (C == 1) {
      // every warp reduces the 32 warp slots itself: one barrier per iteration
      FpsSlot s;
      s.key = static_cast<int>(0x80000000u); s.idx = 0x7fffffff; s.x = s.y = s.z = 0.f;
      if (lane < kFps3Warps) s = wslots[par][lane];
      const int ck = __reduce_max_sync(0xffffffffu, s.key);
      const int cidx = __reduce_min_sync(0xffffffffu, s.key == ck ? s.idx : 0x7fffffff);
      const int src = __ffs(__ballot_sync(0xffffffffu, s.key == ck && s.idx == cidx)) - 1;
      b.idx = cidx;
      b.x = __shfl_sync(0xffffffffu, s.x, src);
      b.y = __shfl_sync(0xffffffffu, s.y, src);
      b.z = __shfl_sync(0xffffffffu, s.z, src);
    } else {
      if (warp == 0) {
        FpsSlot s;
        s.key = static_cast<int>(0x80000000u); s.idx = 0x7fffffff; s.x = s.y = s.z = 0.f;
        if (lane < kFps3Warps) s = wslots[par][lane];
        const int ck = __reduce_max_sync(0xffffffffu, s.key);
        const int cidx = __reduce_min_sync(0xffffffffu, s.key == ck ? s.idx : 0x7fffffff);
        if (s.key == ck && s.idx == cidx) {  // the CTA's winner goes to every CTA of the cluster
          const uint32_t local = smem_u32(&cslots[par][rank]);
          for (int r = 0; r < C; ++r) {
            const uint32_t remote = map_to_cta(local, static_cast<uint32_t>(r));
            st_cluster_v4(remote, s.key, s.idx, __float_as_int(s.x), __float_as_int(s.y));
            st_cluster_v4(remote + 16, __float_as_int(s.z), 0, 0, 0);
          }
        }
      }
      cluster_barrier();
      // every warp picks the cluster winner from the C CTA slots (lane r reads slot r)
      FpsSlot s;
      s.key = static_cast<int>(0x80000000u); s.idx = 0x7fffffff; s.x = s.y = s.z = 0.f;
      if (lane < C) s = cslots[par][lane];
      const int ck = __reduce_max_sync(0xffffffffu, s.key);
      const int cidx = __reduce_min_sync(0xffffffffu, s.key == ck ? s.idx : 0x7fffffff);
      const int src = __ffs(__ballot_sync(0xffffffffu, s.key == ck && s.idx == cidx)) - 1;
      b.idx = cidx;
      b.x = __shfl_sync(0xffffffffu, s.x, src);
      b.y = __shfl_sync(0xffffffffu, s.y, src);
      b.z = __shfl_sync(0xffffffffu, s.z, src);
    }
    lx = b.x; ly = b.y; lz = b.z;
    if (rank == 0 && tid == 0) o[k] = b.idx;
  }
  // keep every CTA's shared memory alive until all remote stores have landed / been read
  if (C > 1) cluster_barrier();
}

// Generic D: one CTA per cloud; mind row in global scratch.
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_generic_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths,
                   const int64_t* __restrict__ Ks, const int64_t* __restrict__ start_idxs, int P,
                   int D, int max_K, float* __restrict__ mind_ws, int64_t* __restrict__ out) {
  __shared__ int wkeys[kFpsWarps], widxs[kFpsWarps];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x;
  int64_t Ll = lengths[n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P ? P : Ll));
  int64_t kl = Ks[n];
  const int kn = static_cast<int>(kl < L ? (kl < 0 ? 0 : kl) : L);
  const float* pts = points + static_cast<size_t>(n) * P * D;
  float* mind = mind_ws + static_cast<size_t>(n) * P;
  int64_t* o = out + static_cast<size_t>(n) * max_K;
  int last = static_cast<int>(start_idxs[n]);
  for (int k = tid; k < max_K; k += kFpsThreads)
    if (k == 0) o[0] = last; else if (k >= kn) o[k] = -1;
  if (kn <= 1) return;
  last = min(max(last, 0), L - 1);
  for (int p = tid; p < L; p += kFpsThreads) mind[p] = FLT_MAX;
  for (int k = 1; k < kn; ++k) {
    const float* lp = pts + static_cast<size_t>(last) * D;
    float best = -1.0f;
    int bi = 0x7fffffff;
    for (int p = tid; p < L; p += kFpsThreads) {
      const float* pp = pts + static_cast<size_t>(p) * D;
      float d = 0.0f;
      for (int dd = 0; dd < D; ++dd) d = __fadd_rn(d, dist_term<2>(lp[dd], pp[dd]));
      const float m = fminf(mind[p], d);
      mind[p] = m;
      if (m > best) { best = m; bi = p; }
    }
    int wkey, widx;
    if (warp_argmax(__float_as_int(best), bi, &wkey, &widx)) { wkeys[warp] = wkey; widxs[warp] = widx; }
    __syncthreads();
    if (warp == 0) {
      int ckey, cidx;
      if (warp_argmax(wkeys[lane], widxs[lane], &ckey, &cidx)) { s_last = cidx; o[k] = cidx; }
    }
    __syncthreads();
    last = s_last;
  }
}

namespace {
template <int PT>
int launch_fps_d3(const float* points, const int64_t* lengths, const int64_t* K,
                  const int64_t* start, int N, int P, int max_K, int C, int64_t* out,
                  cudaStream_t st) {
  auto kern = fps_d3_kernel<PT>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(N) * C);
  cfg.blockDim = dim3(kFps3Threads);
  const int push = get_option("fps_push", 1);  // 1: one-sided mbarrier exchange | 0: CTA winner + cluster barrier
  const size_t smem = (C > 1 && push) ? size_t(3) * PT * kFps3Threads * 4 : 0;
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (C > 8) POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  profile_begin("fps", st);
  POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, points, lengths, K, start, P, max_K, C, push, out));
  profile_end("fps", st);
  POPS_LAUNCH_OK("fps_d3_kernel");
  return POPS_OK;
}
}  // namespace
}  // namespace pops

using namespace pops;

extern "C" size_t pops_fps_workspace_bytes(int64_t N, int64_t P, int64_t D, int64_t max_K) {
  (void)max_K;
  if (D == 3 && P <= int64_t(16) * kFps3Threads * kFps3MaxPT) return 256;
  return align_up(size_t(std::max<int64_t>(N, 0)) * size_t(std::max<int64_t>(P, 0)) * 4, 256) + 256;
}

extern "C" int pops_sample_farthest_points(const float* points, const int64_t* lengths,
                                           const int64_t* K, const int64_t* start_idxs, int64_t N,
                                           int64_t P, int64_t D, int64_t max_K, int64_t* idx,
                                           void* workspace, size_t workspace_bytes,
                                           pops_stream_t stream) {
  POPS_CHECK_ARG(N >= 0 && P >= 0 && D >= 0 && max_K >= 0, "negative size");
  if (N == 0 || max_K == 0) return POPS_OK;
  POPS_CHECK_ARG(points && lengths && K && start_idxs && idx, "null pointer argument");
  POPS_CHECK_ARG(P < (int64_t(1) << 31) && N < (int64_t(1) << 24), "size too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D == 3 && P <= int64_t(16) * kFps3Threads * kFps3MaxPT) {
    // smallest cluster that keeps <= 32 points per thread; widen while the batch leaves SMs idle
    // and threads still have >= 4 points (fewer, fatter threads keep the per-iteration reduction
    // overhead small)
    int C = 1;
    while (int64_t(C) * kFps3Threads * kFps3MaxPT < P) C *= 2;
    const int sms = num_sms();
    while (C < 16 && int64_t(N) * C * 2 <= sms && int64_t(C) * kFps3Threads * 4 < P) C *= 2;
    static const int force_c = getenv("POPS_FPS_C") ? atoi(getenv("POPS_FPS_C")) : 0;  // tuning aid
    if (force_c > 0 && int64_t(force_c) * kFps3Threads * kFps3MaxPT >= P) C = force_c;
    const int per_thread = int(ceil_div(P, int64_t(C) * kFps3Threads));
#define POPS_FPS(PT) return launch_fps_d3<PT>(points, lengths, K, start_idxs, int(N), int(P), int(max_K), C, idx, st)
    if (per_thread <= 1) POPS_FPS(1);
    if (per_thread <= 2) POPS_FPS(2);
    if (per_thread <= 4) POPS_FPS(4);
    if (per_thread <= 8) POPS_FPS(8);
    if (per_thread <= 16) POPS_FPS(16);
    POPS_FPS(32);
#undef POPS_FPS
  }
  if (workspace_bytes < pops_fps_workspace_bytes(N, P, D, max_K) || !workspace)
    return fail(POPS_ERR_WORKSPACE, "fps: workspace missing or too small");
  fps_generic_kernel<<<static_cast<unsigned>(N), kFpsThreads, 0, st>>>(
      points, lengths, K, start_idxs, int(P), int(D), int(max_K), reinterpret_cast<float*>(workspace), idx);
  POPS_LAUNCH_OK("fps_generic_kernel");
  return POPS_OK;
}
