// High-dimensional KNN (32 <= D <= 256, L2, K <= 16) on the 5th-generation tensor cores.
//
// Contract unchanged (knn_cpu.cpp:13-69 of the reference): the K lexicographically smallest
// (dist, idx), dist = the reference's unfused float32 sum.  The tensor cores only FILTER:
//
//   norms    w_j = |y_j|^2 (fp32) for every p2 point, +inf beyond lengths2; max_j w_j per cloud.
//   scan     one CTA = 128 queries of one cloud (the 128 TMEM lanes).  The query tile stays in
//            shared memory; p2 streams through a ring of 16 KB stages (128 points x 32 floats),
//            both moved by TMA straight from the caller's (N,P,D) tensors with the 128-byte
//            swizzle -- no repacking pass.  tcgen05.mma kind::tf32 (one elected thread) accumulates
//            x_i . y_j for 128 x 128 (query, point) pairs into one of four TMEM buffers; four
//            epilogue warps pull the accumulators back with tcgen05.ld and evaluate
//                s_ij = w_j - 2 acc_ij   ~   d(x_i, y_j) - |x_i|^2
//            per pair: one FFMA, one min into a register tournament (32 running minima over
//            disjoint subsets of the columns), and -- only when some lane of the warp has a hit --
//            a predicated append of (s, j) when s <= threshold.  The threshold is
//            (16th smallest of the 32 minima) + 2E: 16 distinct points lie within that minimum, so
//            it bounds the final K-th smallest s from above at any time, and every point the
//            re-rank can need satisfies s <= s_(K) + 2E.  It is refreshed on a geometric schedule.
//            Candidates are staged in shared memory and spilled, re-filtered with the current
//            threshold, to a per-query array in global memory: no sorting, no list upkeep and no
//            p2 row touched outside the tensor cores during the scan.
//   rerank   one warp per query: the 32 smallest (s, j) among the query's candidates (bitonic
//            networks over warp shuffles); with E >= |s_ij + |x_i|^2 - d_ref(i,j)| (bound below)
//            every true neighbour has s <= tau = s_(K) + 2E.  The entries within tau (typically
//            K + a few) get the exact reference distance (same unfused operations, same order),
//            are sorted by the exact 64-bit key, and the first K are the result.
//   fallback if the 32nd smallest s is itself within tau, or a candidate array overflowed, the
//            query is recomputed exactly (knn_exact_rows_kernel; the dense generic kernel when
//            there are thousands): massive ties / duplicate-heavy clouds, adversarial orders.
//
// Error bound.  u = 2^-24.  TF32 operands keep 10 mantissa bits (the low 13 are ignored), so
// |x^y^ - xy| <= (2^-9 + 2^-20)|xy| per product; fp32 accumulation inside the tensor core is
// charged D 2^-22 per unit of sum |x_d y_d| (a 4x allowance over round-to-nearest).  With
// |w~ - |y|^2| <= D u |y|^2, |d_ref - d| <= 2 (D+2) u (|x|^2 + |y|^2), one rounding of the final
// FFMA, and Cauchy-Schwarz:
//     E_i = c1 |x_i| M_y + c2 (|x_i|^2 + M_y^2),  M_y = max_j |y_j|,
//     c1 = 2^-8 (1 + 2^-11 + D 2^-13) + 2^-23,    c2 = (3D + 8) 2^-24,     (+1 % slack)
// The filter can only add candidates; membership and order are decided by exact keys alone.
#include <cuda.h>  // CUtensorMap types; cuTensorMapEncodeTiled is fetched at run time (no -lcuda)

#include <cfloat>
#include <cstdlib>

#include "knn_core.cuh"

namespace pops {

namespace {

constexpr int TC_M = 128;        // queries per CTA = TMEM lanes
constexpr int TC_N = 128;        // points per tile = TMEM columns per accumulator buffer
constexpr int TC_KBLK = 32;      // floats per k-block: one 128-byte swizzle span
constexpr int TC_ABUF = 4;       // accumulator buffers (4 x 128 = all 512 TMEM columns)
constexpr int TC_LIST = 32;      // the re-rank keeps the 32 smallest s of a query (K <= 16)
constexpr int TC_GCAP = 128;     // candidates a (query, column half) can hold in global memory (≈55 used on C5)
constexpr int TC_TOUR = 32;      // tournament minima per query row (TC_HALVES threads x TC_TSLOT); threshold = their 16th smallest
constexpr int TC_TSLOT = 16;     // tournament slots per thread: slot i = minimum over columns 2i, 2i+1 of every 32-column chunk
constexpr int TC_CAND = 24;      // candidate entries a thread stages in shared memory between flushes
constexpr int TC_SUB = 8;        // columns between two staging-overflow checks
constexpr int TC_HALVES = 2;      // epilogue warps per TMEM lane quarter: each takes half of a tile's columns
constexpr int TC_THREADS = 64 + 128 * TC_HALVES;  // warp 0: TMA producer, warp 1: MMA issuer, then the epilogue
constexpr int TC_CSTRIDE = TC_M * TC_HALVES;      // candidate entries between two slots of one buffer
constexpr uint32_t TC_STAGE_BYTES = TC_N * 128;  // 16 KB: 128 rows x 128 bytes
constexpr int TC_MAX_STAGES = 8;

struct TcParams {
  const float* w;          // [N][P2pad]
  const int64_t* len1;
  const int64_t* len2;
  const float* xq;         // [N][P1pad] |x_i|^2
  const unsigned* maxw_bits;  // [N] max_j w_j (float bits)
  uint2* cands;            // [N][P1][TC_HALVES][TC_GCAP] (s bits, j)
  unsigned* counts;        // [N][P1][TC_HALVES] entries used; bit 31 = overflowed
  float* tfin;             // [N][P1] final append threshold: nothing above it can be needed
  int P1, P2, P2pad, P1pad, D;
  int KB;                  // k-blocks = ceil(D / 32)
  int nstage;              // stages of the p2 ring
  int dbg;                 // development: 1 = never buffer a candidate (timing the MMA pipeline alone)
  int seed_tiles;          // tiles of the seed pass (evaluated twice; the real pass starts with their threshold)
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// same load, delivered to the same shared-memory offset of every CTA in `mask` (each CTA's own
// mbarrier at that offset receives the bytes)
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// CTA-pair (cta_group::2) forms.  Shared-memory addresses of a barrier "in the leader" are obtained
// with mapa (rank 0 of the 2-CTA cluster).
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t smem_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(0));
  return r;
}
// TMA load into THIS CTA's shared memory, bytes credited to the barrier at `bar_cluster_addr`
// (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                 uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// Same arrival without release semantics: for "this TMEM buffer is drained", where the only thing the
// waiter depends on is the completion of this warp's tcgen05.ld (tcgen05.wait::ld + fence::before_thread_sync
// precede it).  A release at cluster scope would also wait for every shared / global store of the warp
// (the candidate appends), which nobody on the other side reads.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // one warp of EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A . B^T over the pair: M = 256 (128 rows per CTA), N = 256 (128 rows of B
// from each CTA's shared memory); issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 TMEM lanes (this warp's quarter) x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// Wait for this thread's outstanding tcgen05.ld.  The registers are in/out operands so that no
// use of them can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// K-major, 128-byte-swizzled operand tile (rows of 128 bytes, 8-row atoms 1024 bytes apart):
// start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 >> 4 | version 1 |
// layout SWIZZLE_128B (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
  const uint64_t addr = (smem_u32(smem_tile) >> 4) & 0x3FFFu;
  return addr | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N, M
constexpr uint32_t kTcIdesc =
    (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TC_N >> 3) << 17) | (uint32_t(TC_M >> 4) << 24);
constexpr uint32_t kTcIdescPair =  // cta_group::2: M = 256, N = 256
    (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(256 >> 3) << 17) | (uint32_t(256 >> 4) << 24);

// order-preserving map float -> uint32 (ascending), and back
__device__ __forceinline__ uint32_t f2sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t s) {
  return __uint_as_float(s ^ ((s >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// ---------------------------------------------------------------------------------------------
// norms: one warp per p2 point
// ---------------------------------------------------------------------------------------------
__global__ void tc_norm_kernel(const float* __restrict__ p2, const int64_t* __restrict__ len2, int P2,
                               int P2pad, int D, float* __restrict__ w, unsigned* __restrict__ maxw_bits) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= P2pad) return;
  int64_t Ll = len2[n];
  const int L = static_cast<int>(Ll < 0 ? 0 : (Ll > P2 ? P2 : Ll));
  float acc = 0.0f;
  if (j < L) {
    const float* row = p2 + (static_cast<size_t>(n) * P2 + j) * D;
    for (int d = lane; d < D; d += 32) acc = fmaf(row[d], row[d], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    w[static_cast<size_t>(n) * P2pad + j] = (j < L) ? acc : __int_as_float(0x7f800000);
    if (j < L) atomicMax(maxw_bits + n, __float_as_uint(acc));
  }
}

// ---------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------
// E (see the header) from |x|^2 and max_j |y_j|^2; every factor rounded up, 1 % slack on top.
// The scan and the re-rank call this with identical arguments, so they agree bit for bit.
__device__ __forceinline__ float tc_error_bound(float xx, float maxw, int D) {
  const float My2 = maxw * 1.0001f;
  const float xx_up = xx * 1.0001f;
  const float c1 = 0.00390625f * (1.0f + 0.00048828125f + static_cast<float>(D) * 0.0001220703125f) + 1.2e-7f;
  const float c2 = static_cast<float>(3 * D + 8) * 5.9604645e-08f;
  return 1.01f * (c1 * sqrtf(xx_up * My2) * 1.0001f + c2 * (xx_up + My2));
}

// Spill one thread's staged candidates (column stride TC_CSTRIDE) to its array in global memory,
// keeping only those still within the current threshold.  When the array is full it is first
// compacted against the threshold; if that does not help the overflow bit is set and the query
// will be recomputed exactly.  Not inlined: rare.  Returns the new count.
__device__ __noinline__ unsigned tc_flush_stage(const uint2* cand_col, int staged, float T, uint2* garr,
                                                unsigned gcount) {
  for (int c = 0; c < staged; ++c) {
    const uint2 e = cand_col[c * TC_CSTRIDE];
    if (!(__uint_as_float(e.x) <= T)) continue;
    unsigned cnt = gcount & 0x7fffffffu;
    if (cnt == TC_GCAP) {
      unsigned wpos = 0;
      for (unsigned r = 0; r < TC_GCAP; ++r) {
        const uint2 g = garr[r];
        if (__uint_as_float(g.x) <= T) garr[wpos++] = g;
      }
      cnt = wpos;
      gcount = (gcount & 0x80000000u) | cnt;
      if (cnt > TC_GCAP - 32) {  // still (nearly) full: give up on this query
        gcount |= 0x80000000u;
        if (cnt == TC_GCAP) continue;
      }
    }
    garr[cnt] = e;
    gcount = (gcount & 0x80000000u) | (cnt + 1);
  }
  return gcount;
}

// CL = CTAs per cluster.  The CTAs of a cluster work on CL consecutive query tiles of ONE cloud and
// walk the same p2 tiles in lock step: each loads 1/CL of every stage and multicasts it to all, so
// L2 -> shared-memory traffic drops by CL (one CTA per 128 queries streaming all of p2 by itself
// needs more than the L2 can deliver: measured 5.1 TB/s, 3x the MMA time).
// PAIR: the two CTAs of a cluster form a cta_group::2 pair.  One tcgen05.mma of the leader covers
// 256 queries x 256 points: every CTA supplies its 128 query rows and 128 of the 256 p2 rows of a
// tile from its own shared memory and receives the accumulators of ITS queries (128 lanes x 256
// columns).  Per SM the MMA then reads 64 B/clk of shared memory instead of 128 and TMA writes 32
// instead of 64 -- the single-CTA form is bound by exactly that bandwidth.
template <int CL, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_scan_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_p,
                   const TcParams prm) {
  extern __shared__ unsigned char smem_raw[];
  // the 128-byte swizzle needs 1024-byte aligned tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr unsigned FULL = 0xffffffffu;
  const int n = blockIdx.y;
  const int q_base = blockIdx.x * TC_M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  // the rerank kernel writes the (0, 0) rows.  A CTA without valid queries still has to feed and
  // release its cluster (L2 is the same for the whole cluster: same cloud).
  static_assert(!PAIR || CL == 1, "the pair shares p2 through the MMA, not through multicast");
  if (L2 == 0 || (CL == 1 && !PAIR && q_base >= L1)) return;
  const uint32_t crank = (CL > 1 || PAIR) ? cluster_ctarank() : 0;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);
  constexpr int TN = PAIR ? 2 * TC_N : TC_N;   // accumulator columns (= points) per tile
  constexpr int ABUF = 512 / TN;               // accumulator buffers: all 512 TMEM columns
  const bool leader = crank == 0;
  const int KB = prm.KB, NST = prm.nstage;
  const int num_tiles = (L2 + TN - 1) / TN;
  // Seed pass: the first tiles are evaluated twice -- first for the tournament only (no candidate
  // is staged), so that the real pass starts with a finite threshold instead of buffering the first
  // few hundred points of every query.  Sequence of tiles for all roles: 0..S-1, then 0..num_tiles-1.
  const int S_seed = num_tiles >= 4 * prm.seed_tiles ? prm.seed_tiles : 0;
  const int seq_tiles = S_seed + num_tiles;

  float* sA = reinterpret_cast<float*>(smem);                                  // KB x 16 KB
  float* sB = reinterpret_cast<float*>(smem + static_cast<size_t>(KB) * TC_STAGE_BYTES);  // NST x 16 KB
  unsigned char* rest = smem + static_cast<size_t>(KB + NST) * TC_STAGE_BYTES;
  uint2* cand = reinterpret_cast<uint2*>(rest);                                // TC_CAND x TC_CSTRIDE staged (s, j)
  float* sX = reinterpret_cast<float*>(rest + size_t(TC_CAND) * TC_CSTRIDE * 8);  // TC_TSLOT x TC_M: minima exchange
  float* sW = sX + TC_TSLOT * TC_M;                                             // per epilogue warp: 2 x TC_N/TC_HALVES norms
  float* sT = sW + 4 * TC_HALVES * 2 * (2 * TC_N / TC_HALVES);                 // TC_M tournament bounds
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + TC_M);
  uint64_t* full = bars;                       // [TC_MAX_STAGES]  TMA -> MMA
  uint64_t* empty = bars + TC_MAX_STAGES;      // [TC_MAX_STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * TC_MAX_STAGES;  // [TC_ABUF]        MMA -> epilogue
  uint64_t* tempty = tfull + TC_ABUF;          // [TC_ABUF]        epilogue -> MMA
  uint64_t* afull = tempty + TC_ABUF;          // query tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);  // every CTA of the cluster has consumed the stage
    }
    for (int b = 0; b < ABUF; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4 * TC_HALVES * (PAIR ? 2 : 1));  // one arrival per epilogue warp (of both CTAs)
    }
    mbar_init(afull, 1);
    mbar_fence_init();
  } else if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1 || PAIR) cluster_sync_all();  // peers' barriers are initialised before anything arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one thread) =====
    if (lane == 0) {
      // PAIR: the MMA issuer lives in the leader, so every load of either CTA is credited to the
      // LEADER's barriers, which expect the bytes of both
      if (!PAIR || leader) mbar_arrive_expect_tx(afull, static_cast<uint32_t>(KB) * TC_STAGE_BYTES * (PAIR ? 2 : 1));
      const uint32_t afull_l = PAIR ? mapa_rank0(smem_u32(afull)) : 0;
      for (int kb = 0; kb < KB; ++kb) {
        if (PAIR)
          tma_load_3d_pair(sA + static_cast<size_t>(kb) * (TC_STAGE_BYTES / 4), &map_q, kb * TC_KBLK, q_base, n, afull_l);
        else
          tma_load_3d(sA + static_cast<size_t>(kb) * (TC_STAGE_BYTES / 4), &map_q, kb * TC_KBLK, q_base, n, afull);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int tt = 0; tt < seq_tiles; ++tt) {
        const int t = tt < S_seed ? tt : tt - S_seed;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_parked(&empty[s], ph ^ 1);
          if (PAIR) {  // this CTA's 128 of the tile's 256 rows
            if (leader) mbar_arrive_expect_tx(&full[s], 2 * TC_STAGE_BYTES);
            tma_load_3d_pair(sB + static_cast<size_t>(s) * (TC_STAGE_BYTES / 4), &map_p, kb * TC_KBLK,
                             t * TN + static_cast<int>(crank) * TC_N, n, mapa_rank0(smem_u32(&full[s])));
            if (++s == NST) {
              s = 0;
              ph ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
          if (CL == 1) {
            tma_load_3d(sB + static_cast<size_t>(s) * (TC_STAGE_BYTES / 4), &map_p, kb * TC_KBLK, t * TC_N, n, &full[s]);
          } else {  // this CTA's slice of the rows, to every CTA of the cluster
            constexpr int SL = TC_N / CL;
            tma_load_3d_mc(sB + static_cast<size_t>(s) * (TC_STAGE_BYTES / 4) + crank * (SL * 32), &map_p, kb * TC_KBLK,
                           t * TC_N + static_cast<int>(crank) * SL, n, &full[s], kMask);
          }
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && (!PAIR || leader)) {
      mbar_wait(afull, 0);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < seq_tiles; ++t) {  // t: position in the sequence (buffers rotate with it)
        const int b = t % ABUF;
        mbar_wait_parked(&tempty[b], ((t / ABUF) & 1) ^ 1);  // the epilogue(s) drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * TN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_parked(&full[s], ph);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(sA + static_cast<size_t>(kb) * (TC_STAGE_BYTES / 4));
          const uint64_t bdesc = umma_desc_sw128(sB + static_cast<size_t>(s) * (TC_STAGE_BYTES / 4));
#pragma unroll
          for (int k = 0; k < TC_KBLK / 8; ++k) {  // K = 8 per tf32 MMA: 32 bytes along the swizzled row
            if (PAIR)
              umma_tf32_pair(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2),
                             kTcIdescPair, (kb | k) != 0 ? 1u : 0u);
            else
              umma_tf32(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), kTcIdesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          // the stage may be refilled once these MMAs have read it
          if (PAIR) umma_commit_pair(&empty[s]);
          else if (CL == 1) umma_commit(&empty[s]);
          else umma_commit_mc(&empty[s], kMask);
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
        }
        if (PAIR) umma_commit_pair(&tfull[b]); else umma_commit(&tfull[b]);
      }
    }
  } else {
    // ===== epilogue: 4 x TC_HALVES warps.  Thread = (query row, half of every tile's columns) =====
    const int ew = warp & 3;                 // the TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;        // which half of the columns
    const int row = ew * 32 + lane;          // query row inside the tile
    const int qi = q_base + row;
    const float INF = __int_as_float(0x7f800000);
    constexpr int HC = TN / TC_HALVES;       // columns per half
    static_assert(HC % 32 == 0 && TC_TSLOT == 16 && TC_TOUR == TC_HALVES * TC_TSLOT && TC_SUB == 8,
                  "columns 2i, 2i+1 of a chunk feed tournament slot i; the two threads of a row hold 32 disjoint subsets");
    const bool live = qi < L1 && !(prm.dbg & 1);
    const size_t qrow = static_cast<size_t>(n) * prm.P1 + min(qi, prm.P1 - 1);
    const float E2 = 2.0f * tc_error_bound(prm.xq[static_cast<size_t>(n) * prm.P1pad + min(qi, prm.P1pad - 1)],
                                           __uint_as_float(prm.maxw_bits[n]), prm.D);
    float mins[TC_TSLOT];  // running minima of s over 16 disjoint subsets of this thread's columns
#pragma unroll
    for (int i = 0; i < TC_TSLOT; ++i) mins[i] = INF;
    float T = live ? INF : -INF;   // append threshold; rows beyond lengths1 never buffer anything
    uint2* garr = prm.cands + (qrow * TC_HALVES + half) * TC_GCAP;
    unsigned gcount = 0;
    uint2* cand_col = cand + half * TC_M + row;
    const uint32_t cand_base = smem_u32(cand_col);
    uint32_t cw = cand_base;
    constexpr uint32_t CSTRIDE = TC_CSTRIDE * 8;
    const uint32_t cw_limit = cand_base + static_cast<uint32_t>(TC_CAND - TC_SUB) * CSTRIDE;
    // every warp stages the norms of ITS columns of the next tile in a private double buffer: no
    // CTA-wide barrier per tile
    constexpr int WPL = HC / 32;  // norms per lane: 2 or 4
    const float* w_n = prm.w + static_cast<size_t>(n) * prm.P2pad + half * HC;
    float* sWw = sW + (warp - 2) * 2 * HC;
    const uint32_t sW_addr = smem_u32(sWw);
    const uint32_t tempty_l = PAIR ? mapa_rank0(smem_u32(tempty)) : 0;  // the leader's tempty[0]
#pragma unroll
    for (int i = 0; i < WPL; ++i) sWw[WPL * lane + i] = w_n[WPL * lane + i];
    __syncwarp();
    for (int tt = 0; tt < seq_tiles; ++tt) {
      const bool seeding = tt < S_seed;
      const int t = seeding ? tt : tt - S_seed;                                // the tile
      const int tn = (tt + 1 < S_seed) ? tt + 1 : tt + 1 - S_seed;             // the next tile of the sequence
      const int b = tt % ABUF;
      const uint32_t wt = sW_addr + static_cast<uint32_t>((tt & 1) * HC * 4);
      float wnext[WPL];
#pragma unroll
      for (int i = 0; i < WPL; ++i) wnext[i] = (tt + 1 < seq_tiles) ? w_n[tn * TN + WPL * lane + i] : 0.0f;
      mbar_wait_parked(&tfull[b], (tt / ABUF) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(b * TN + half * HC);
      // The accumulators of chunk c + 1 are requested as soon as chunk c's have been turned into s values
      // (its registers are free then) and arrive while the hit tests and appends of chunk c run.
      const int nchunk = (prm.dbg & 2) ? 0 : HC / 32;  // dbg 2: drain nothing (MMA / TMA pipeline alone)
      uint32_t acc[32];
      if (nchunk > 0) tmem_ld_x32(taddr, acc);
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        tmem_ld_wait(acc);
        if (prm.dbg & 4) {  // dbg 4: read the accumulators, evaluate nothing
          if (acc[0] == 0x12345678u && acc[31] == 0x9abcdef0u) T = 0.0f;
          if (c + 1 < nchunk) tmem_ld_x32(taddr + static_cast<uint32_t>((c + 1) * 32), acc);
          continue;
        }
        const uint32_t jc = static_cast<uint32_t>(t * TN + half * HC + c * 32);
        // straight-line part first (32 independent columns: norms, s, pair minima, tournament), hit tests after
        float sv[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float w0, w1, w2, w3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(w0), "=f"(w1), "=f"(w2), "=f"(w3)
                       : "r"(wt + static_cast<uint32_t>((c * 32 + i4 * 4) * 4)));
          sv[i4 * 4 + 0] = fmaf(-2.0f, __uint_as_float(acc[i4 * 4 + 0]), w0);
          sv[i4 * 4 + 1] = fmaf(-2.0f, __uint_as_float(acc[i4 * 4 + 1]), w1);
          sv[i4 * 4 + 2] = fmaf(-2.0f, __uint_as_float(acc[i4 * 4 + 2]), w2);
          sv[i4 * 4 + 3] = fmaf(-2.0f, __uint_as_float(acc[i4 * 4 + 3]), w3);
        }
        if (c + 1 < nchunk) tmem_ld_x32(taddr + static_cast<uint32_t>((c + 1) * 32), acc);
        float m2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          m2[i] = fminf(sv[2 * i], sv[2 * i + 1]);
          mins[i] = fminf(mins[i], m2[i]);
        }
        float m8[32 / TC_SUB];
#pragma unroll
        for (int sub = 0; sub < 32 / TC_SUB; ++sub)
          m8[sub] = fminf(fminf(m2[sub * 4], m2[sub * 4 + 1]), fminf(m2[sub * 4 + 2], m2[sub * 4 + 3]));
        // the append code is skipped unless some lane has a hit among the columns (never while seeding)
        if (!seeding && __any_sync(FULL, fminf(fminf(m8[0], m8[1]), fminf(m8[2], m8[3])) <= T)) {
#pragma unroll
          for (int sub = 0; sub < 32 / TC_SUB; ++sub) {
            if (!__any_sync(FULL, m8[sub] <= T)) continue;
#pragma unroll
            for (int i = 0; i < TC_SUB; ++i) {
              if (sv[sub * TC_SUB + i] <= T) {  // predicated: one 64-bit store + one add (no "memory" clobber:
                                                // the buffer is only read back inside tc_flush_stage, an opaque call)
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(cw), "r"(__float_as_uint(sv[sub * TC_SUB + i])),
                             "r"(jc + static_cast<uint32_t>(sub * TC_SUB + i)));
                cw += CSTRIDE;
              }
            }
            if (__any_sync(FULL, cw > cw_limit)) {
              // a real call: the accumulators in flight must have landed before their registers may be
              // saved and restored around it
              tmem_ld_wait(acc);
              gcount = tc_flush_stage(cand_col, static_cast<int>((cw - cand_base) / CSTRIDE), T, garr, gcount);
              cw = cand_base;
            }
          }
        }
      }
      // accumulator buffer drained
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster_relaxed(tempty_l + static_cast<uint32_t>(b * 8)); else mbar_arrive(&tempty[b]);
      }
      // refresh the threshold: after tiles 1, 2, 3, 4, 6, 8, 12, 16, ... and then every 16th.  The
      // two threads of a row (one per column half) hold 16 minima each over disjoint subsets of their
      // columns; half 0 takes the 16th smallest of the 32.
      const int v = t + 1;
      const bool pow2 = (v & (v - 1)) == 0;
      const bool pow2x3 = (v % 3 == 0) && (((v / 3) & ((v / 3) - 1)) == 0);
      const bool refresh = seeding ? (tt == S_seed - 1) : (v < 64 ? (pow2 || pow2x3) : (v & 15) == 0);
      if (refresh) {
        if (half != 0) {
#pragma unroll
          for (int i = 0; i < TC_TSLOT; ++i) sX[i * TC_M + row] = mins[i];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_HALVES) : "memory");
        if (half == 0) {
          // 16th smallest of the row's 32 minima (16 of each half's thread, 32 disjoint subsets of everything
          // the row has seen): sort both halves, then max_i min(A[i], B[15-i])
          float tmp[TC_TOUR];
#pragma unroll
          for (int i = 0; i < TC_TSLOT; ++i) {
            tmp[i] = mins[i];
            tmp[TC_TSLOT + i] = sX[i * TC_M + row];
          }
          sort_floats<16, 0, TC_TOUR>(tmp);
          sort_floats<16, 16, TC_TOUR>(tmp);
          float U = fminf(tmp[0], tmp[31]);
#pragma unroll
          for (int i = 1; i < 16; ++i) U = fmaxf(U, fminf(tmp[i], tmp[31 - i]));
          sT[row] = U;
        }
      }
      // stage the next tile's norms (the warp is past its reads of that buffer)
#pragma unroll
      for (int i = 0; i < WPL; ++i) sWw[((tt + 1) & 1) * HC + WPL * lane + i] = wnext[i];
      __syncwarp();
      if (refresh) {
        asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_HALVES) : "memory");
        if (live) T = __fadd_ru(sT[row], E2);  // +inf while fewer than 16 points have been seen
      }
    }
    gcount = tc_flush_stage(cand_col, static_cast<int>((cw - cand_base) / CSTRIDE), T, garr, gcount);
    if (qi < L1) {
      prm.counts[qrow * TC_HALVES + half] = gcount;
      if (half == 0) prm.tfin[qrow] = T;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1 || PAIR) cluster_sync_all();  // nobody leaves while a peer may still write or signal here
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// rerank: one warp per query
// ---------------------------------------------------------------------------------------------
__device__ unsigned g_tc_dbg = 0;

struct TcRerankParams {
  const float* p1;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  const uint2* cands;     // [N][P1][TC_HALVES][TC_GCAP]
  const unsigned* counts; // [N][P1][TC_HALVES]
  const float* tfin;      // [N][P1]
  const float* xq;        // [N][P1pad]
  const unsigned* maxw_bits;
  const unsigned* maxq_bits;  // max |x|^2 per cloud (bits), like maxw_bits for p2
  int64_t* idx;
  float* dists;
  unsigned char* flags;  // [N][P1]: 1 = recompute exactly
  unsigned* flag_rows;   // compact list of the flagged rows (n * P1 + i); flag_rows[-1] is the counter
  int P1, P2, P1pad, D, K;
  int debug;
};

__global__ void __launch_bounds__(128) knn_tc_rerank_kernel(const TcRerankParams prm) {
  extern __shared__ float sx[];  // [4][D]: the warps' query rows
  constexpr unsigned FULL = 0xffffffffu;
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int qi = blockIdx.x * 4 + wib;
  if (qi >= prm.P1) return;
  const int D = prm.D, K = prm.K;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  const size_t qrow = static_cast<size_t>(n) * prm.P1 + qi;
  int64_t* oi = prm.idx + qrow * K;
  float* od = prm.dists + qrow * K;
  if (lane == 0) prm.flags[qrow] = 0;
  if (qi >= L1 || L2 == 0) {
    for (int k = lane; k < K; k += 32) {
      oi[k] = 0;
      od[k] = 0.0f;
    }
    return;
  }
  // query row -> shared memory
  float* x = sx + wib * D;
  const float* xg = prm.p1 + qrow * D;
  for (int d = lane; d < D; d += 32) x[d] = xg[d];
  __syncwarp();
  const float E = tc_error_bound(prm.xq[static_cast<size_t>(n) * prm.P1pad + qi], __uint_as_float(prm.maxw_bits[n]), D);
  // a squared norm that is +inf, NaN or beyond 1e36 voids the filter's error bound for the whole cloud
  // (common.cuh): every one of its queries goes to the exact recomputation
  const bool dirty = prm.maxw_bits[n] >= kDirtyNormBits || prm.maxq_bits[n] >= kDirtyNormBits;

  // the 32 smallest (s, j) among the query's candidates, ascending along the lanes: every chunk of
  // 32 candidates is sorted (bitonic network over shuffles); min(run[i], chunk[31-i]) holds the 32
  // smallest of both as a bitonic sequence, which one merge pass sorts
  auto sort32 = [&](uint64_t v) -> uint64_t {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const uint64_t other = __shfl_xor_sync(FULL, v, stride);
        const bool up = (lane & size) == 0;
        const bool lower = (lane & stride) == 0;
        v = (up == lower) ? (other < v ? other : v) : (other > v ? other : v);
      }
    }
    return v;
  };
  auto merge32 = [&](uint64_t v) -> uint64_t {  // v bitonic -> ascending
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
      const uint64_t other = __shfl_xor_sync(FULL, v, stride);
      v = (lane & stride) == 0 ? (other < v ? other : v) : (other > v ? other : v);
    }
    return v;
  };
  // Only candidates within the scan's FINAL threshold can matter (tau <= that threshold); they are
  // compacted through a small per-warp queue so that the sorting networks run on full chunks.
  __shared__ uint64_t qbuf_all[4][64];
  uint64_t* qbuf = qbuf_all[wib];
  const float Tfin = prm.tfin[qrow];
  const unsigned lt_mask = (1u << lane) - 1u;
  uint64_t run = kEmptyKey;
  bool unsure = dirty;
  unsigned total = 0, fill = 0;
  for (int h = 0; h < (dirty ? 0 : TC_HALVES); ++h) {
    const unsigned raw = prm.counts[qrow * TC_HALVES + h];
    unsure = unsure || (raw >> 31) != 0;  // the candidate array overflowed
    const unsigned cnt = raw & 0x7fffffffu;
    const uint2* arr = prm.cands + (qrow * TC_HALVES + h) * TC_GCAP;
    for (unsigned c0 = 0; c0 < cnt; c0 += 32) {
      bool pass = false;
      uint64_t key = kEmptyKey;
      if (c0 + lane < cnt) {
        const uint2 e = arr[c0 + lane];
        pass = __uint_as_float(e.x) <= Tfin;
        key = (static_cast<uint64_t>(f2sortable(__uint_as_float(e.x))) << 32) | e.y;
      }
      const unsigned m = __ballot_sync(FULL, pass);
      if (pass) qbuf[fill + __popc(m & lt_mask)] = key;
      fill += __popc(m);
      __syncwarp();
      if (fill >= 32) {
        uint64_t k = sort32(qbuf[lane]);
        const uint64_t rev = __shfl_sync(FULL, k, 31 - lane);
        run = merge32(rev < run ? rev : run);
        const uint64_t tail = qbuf[32 + lane];
        __syncwarp();
        qbuf[lane] = tail;
        fill -= 32;
        total += 32;
        __syncwarp();
      }
    }
  }
  if (fill > 0) {
    uint64_t k = sort32(lane < fill ? qbuf[lane] : kEmptyKey);
    const uint64_t rev = __shfl_sync(FULL, k, 31 - lane);
    run = merge32(rev < run ? rev : run);
    total += fill;
  }
  const bool have = run != kEmptyKey;
  const float s = have ? sortable2f(static_cast<uint32_t>(run >> 32)) : __int_as_float(0x7f800000);
  const uint32_t j = static_cast<uint32_t>(run & 0xFFFFFFFFull);
  const int kth = min(K, L2);
  const float sK = __shfl_sync(FULL, s, kth - 1);  // lane k holds the (k+1)-th smallest s
  const float tau = __fadd_ru(sK, __fmul_ru(2.0f, E));
  {
    const float s_last = __shfl_sync(FULL, s, TC_LIST - 1);
    // more than 32 candidates and the 32nd is still within tau: the band is not fully visible here
    unsure = unsure || (total > TC_LIST && s_last <= tau);
  }
  if (unsure) {
    if (lane == 0) {
      prm.flags[qrow] = 1;
      prm.flag_rows[atomicAdd(prm.flag_rows - 1, 1u)] = static_cast<unsigned>(qrow);
    }
    if (prm.debug && lane == 0 && atomicAdd(&g_tc_dbg, 1u) < 12u)
      printf("flag n=%d q=%d sK=%g tau=%g E=%g total=%u kth=%d\n", n, qi, sK, tau, E, total, kth);
    return;
  }
  if (prm.debug && lane == 0 && (qrow & 0xFFFF) == 0)
    printf("row %d: %u candidates within the final threshold of %u + %u appended\n", static_cast<int>(qrow), total,
           prm.counts[qrow * TC_HALVES] & 0x7fffffffu, prm.counts[qrow * TC_HALVES + 1] & 0x7fffffffu);
  // exact reference distance for the entries within tau: same operations, same order (knn_cpu.cpp:42-50)
  uint64_t k2 = kEmptyKey;
  if (have && s <= tau) {
    const float* y = prm.p2 + (static_cast<size_t>(n) * prm.P2 + j) * D;
    float dist = 0.0f;
    if ((D & 3) == 0) {
      for (int d = 0; d < D; d += 4) {
        const float4 yv = *reinterpret_cast<const float4*>(y + d);
        const float4 xv = *reinterpret_cast<const float4*>(x + d);
        float df = __fsub_rn(xv.x, yv.x);
        dist = __fadd_rn(dist, __fmul_rn(df, df));
        df = __fsub_rn(xv.y, yv.y);
        dist = __fadd_rn(dist, __fmul_rn(df, df));
        df = __fsub_rn(xv.z, yv.z);
        dist = __fadd_rn(dist, __fmul_rn(df, df));
        df = __fsub_rn(xv.w, yv.w);
        dist = __fadd_rn(dist, __fmul_rn(df, df));
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const float df = __fsub_rn(x[d], y[d]);
        dist = __fadd_rn(dist, __fmul_rn(df, df));
      }
    }
    k2 = make_key(dist, j);
  }
  k2 = sort32(k2);
  if (lane < K) {
    const bool ok = k2 != kEmptyKey;
    oi[lane] = ok ? static_cast<int64_t>(k2 & 0xFFFFFFFFull) : 0;
    od[lane] = ok ? key_dist(k2) : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// exact recomputation of the flagged queries: one CLUSTER of kExactCluster CTAs per flagged row.  Every
// CTA scans its share of p2 with the reference arithmetic (each thread keeps its own sorted top-KT in
// registers; K rounds of block-wide arg-min over the threads' current heads give the CTA's K best in
// order), the CTAs' K-lists meet in the shared memory of CTA 0 of the cluster (DSMEM reads after one
// cluster barrier), and one warp picks the K smallest of them.  One CTA per row spent 0.55 ms on 11 rows
// of C5 (a 16 MB cloud streamed by one SM).
// Rows are few on generic data (a handful per million queries); beyond `limit` rows the dense
// generic kernel (knn.cu) takes over instead, CTA by CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kExactCluster = 8;
constexpr int kExactThreads = 1024;

template <int KT>
__global__ void __launch_bounds__(kExactThreads)
knn_exact_rows_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                      const int64_t* __restrict__ len2, const unsigned* __restrict__ flag_rows, unsigned limit,
                      int P1, int P2, int D, int K, int64_t* __restrict__ idx, float* __restrict__ dists) {
  extern __shared__ __align__(16) unsigned char esm[];
  float* x = reinterpret_cast<float*>(esm);                                   // [D]
  uint64_t* red = reinterpret_cast<uint64_t*>(esm + align_up(size_t(D) * 4, 16));  // [32] warp minima + [1] winner + [16] this CTA's K best
  uint64_t* loc = red + 34;
  float* tile = reinterpret_cast<float*>(red + 50) + (threadIdx.x >> 5) * (32 * 33);       // per warp: 32 x 33
  const unsigned count = flag_rows[-1];
  if (count > limit) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned crank = cluster_ctarank();
  const unsigned ncl = gridDim.x / kExactCluster;
  static_assert(kExactCluster * 16 <= 32 * 4, "the merging warp holds 4 keys per lane");
  // every CTA of a cluster runs the same number of trips: the cluster barriers inside line up
  for (unsigned f = blockIdx.x / kExactCluster; f < count; f += ncl) {
    const size_t qrow = flag_rows[f];
    const int n = static_cast<int>(qrow / P1);
    int64_t L2l = len2[n];
    const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > P2 ? P2 : L2l));
    __syncthreads();
    for (int d = tid; d < D; d += kExactThreads) x[d] = p1[qrow * D + d];
    __syncthreads();
    uint64_t Lr[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) Lr[k] = kEmptyKey;
    // a warp takes 32 points at a time: rows are read coalesced into a padded tile, then every
    // lane sums its own point in dimension order (the reference's order)
    const float* yb = p2 + static_cast<size_t>(n) * P2 * D;
    for (int base = (static_cast<int>(crank) * (kExactThreads / 32) + warp) * 32; base < L2;
         base += kExactCluster * (kExactThreads / 32) * 32) {
      float dist = 0.0f;
      for (int d0 = 0; d0 < D; d0 += 32) {
        const int dn = min(32, D - d0);
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {  // 32 independent coalesced loads in flight
          v[r] = 0.0f;
          if (base + r < L2 && lane < dn) v[r] = yb[static_cast<size_t>(base + r) * D + d0 + lane];
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r * 33 + lane] = v[r];
        __syncwarp();
        for (int dd = 0; dd < dn; ++dd) {
          const float df = __fsub_rn(x[d0 + dd], tile[lane * 33 + dd]);
          dist = __fadd_rn(dist, __fmul_rn(df, df));
        }
        __syncwarp();
      }
      const int j = base + lane;
      if (j < L2) {
        const uint64_t key = make_key_total(dist, static_cast<uint32_t>(j));  // finite < +inf < NaN (common.cuh)
        if (key < Lr[KT - 1]) insert_network<KT>(Lr, key);
      }
    }
    // K rounds: smallest head wins and its owner pops
    for (int k = 0; k < K; ++k) {
      uint64_t m = Lr[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v < m ? v : m;
      }
      if (lane == 0) red[warp] = m;
      __syncthreads();
      if (warp == 0) {
        uint64_t v = lane < kExactThreads / 32 ? red[lane] : kEmptyKey;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t u = __shfl_xor_sync(0xffffffffu, v, o);
          v = u < v ? u : v;
        }
        if (lane == 0) red[32] = v;
      }
      __syncthreads();
      const uint64_t win = red[32];
      if (win != kEmptyKey && Lr[0] == win) {  // keys are unique: exactly one owner
#pragma unroll
        for (int q = 0; q + 1 < KT; ++q) Lr[q] = Lr[q + 1];
        Lr[KT - 1] = kEmptyKey;
      }
      if (tid == 0) loc[k] = win;
      __syncthreads();
    }
    // the cluster's lists -> CTA 0, warp 0: lane e holds keys e, e + 32, ... of the kExactCluster * K
    cluster_sync_all();
    if (crank == 0 && warp == 0) {
      uint64_t mine[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = lane + 32 * r;
        mine[r] = kEmptyKey;
        if (e < kExactCluster * K) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(loc + e % K)), "r"(e / K));
          asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(mine[r]) : "r"(ra) : "memory");
        }
      }
      for (int k = 0; k < K; ++k) {
        uint64_t m = mine[0];
#pragma unroll
        for (int r = 1; r < 4; ++r) m = mine[r] < m ? mine[r] : m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t v = __shfl_xor_sync(0xffffffffu, m, o);
          m = v < m ? v : m;
        }
        if (m != kEmptyKey) {  // keys are unique: exactly one owner drops it
#pragma unroll
          for (int r = 0; r < 4; ++r) mine[r] = mine[r] == m ? kEmptyKey : mine[r];
        }
        if (lane == 0) {
          idx[qrow * K + k] = m != kEmptyKey ? static_cast<int64_t>(m & 0xFFFFFFFFull) : 0;
          dists[qrow * K + k] = m != kEmptyKey ? key_dist(m) : 0.0f;
        }
      }
    }
    cluster_sync_all();  // the peers' lists stay alive until CTA 0 has read them
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (N, P, D) f32 row-major as a 3-D tensor map: box = 32 floats x `rows` points x 1 cloud, 128B swizzle
int make_map(CUtensorMap* map, const float* base, int64_t N, int64_t P, int64_t D, int rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(POPS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {cuuint64_t(D), cuuint64_t(P), cuuint64_t(N)};
  const cuuint64_t strides[2] = {cuuint64_t(D) * 4, cuuint64_t(P) * cuuint64_t(D) * 4};
  const cuuint32_t box[3] = {TC_KBLK, cuuint32_t(rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(POPS_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
  return POPS_OK;
}

struct TcLayout {
  int P2pad;
  int P1pad;
  size_t w_off, xq_off, maxw_off, maxq_off, cands_off, counts_off, tfin_off, flags_off, rows_off, total;
};

TcLayout tc_layout(int64_t N, int64_t P1, int64_t P2) {
  TcLayout l;
  l.P2pad = static_cast<int>((P2 + 2 * TC_N - 1) / (2 * TC_N) * (2 * TC_N));  // whole 256-point tiles
  size_t off = 0;
  l.P1pad = static_cast<int>((P1 + TC_M - 1) / TC_M * TC_M);
  l.maxw_off = off;  off += align_up(size_t(N) * 4, 256);
  l.maxq_off = off;  off += align_up(size_t(N) * 4, 256);
  l.w_off = off;     off += align_up(size_t(N) * l.P2pad * 4, 256);
  l.xq_off = off;    off += align_up(size_t(N) * l.P1pad * 4, 256);
  l.cands_off = off; off += align_up(size_t(N) * P1 * TC_HALVES * TC_GCAP * 8, 256);
  l.counts_off = off; off += align_up(size_t(N) * P1 * TC_HALVES * 4, 256);
  l.tfin_off = off;  off += align_up(size_t(N) * P1 * 4, 256);
  l.flags_off = off; off += align_up(size_t(N) * P1, 256);
  l.rows_off = off;  off += align_up(size_t(N) * P1 * 4 + 256, 256);  // counter lives 4 bytes before the list
  l.total = off;
  return l;
}

constexpr size_t kSmemLimit = 227 * 1024;
inline size_t tc_smem_fixed(int KB) {
  return size_t(KB) * TC_STAGE_BYTES + size_t(TC_CAND) * TC_CSTRIDE * 8 + size_t(TC_TSLOT) * TC_M * 4 + 4 * 2 * 2 * TC_N * 4 + TC_M * 4 +
         (2 * TC_MAX_STAGES + 2 * TC_ABUF + 2) * 8 +
         1024 /* alignment */;
}

}  // namespace

bool knn_tc_supported(int64_t P1, int64_t P2, int64_t D, int64_t K, int norm) {
  const int force = get_option("knn_tc", -1);  // test aid: 0 = never, 1 = whenever the shape allows
  if (force == 0) return false;
  if (norm != 2 || K > 16 || K < 1 || D < 32 || D > 256 || (D & 3) != 0) return false;
  if (P2 >= (int64_t(1) << 31) - TC_N) return false;
  const int KB = static_cast<int>((D + TC_KBLK - 1) / TC_KBLK);
  if (tc_smem_fixed(KB) + 2 * TC_STAGE_BYTES > kSmemLimit) return false;
  if (force > 0) return true;
  return P2 >= 512 && P1 >= 64;  // below that the exact generic kernel is as fast
}

size_t knn_tc_workspace_bytes(int64_t N, int64_t P1, int64_t P2) { return tc_layout(N, P1, P2).total + 256; }

// Runs norms + scan + rerank; on return `*flags_out` (N*P1 bytes, device) marks the queries the
// caller must recompute with the exact generic kernel.
int knn_tc_search(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N, int P1,
                  int P2, int D, int K, int64_t* idx, float* dists, void* ws, unsigned char** flags_out,
                  const unsigned** flag_count_out, unsigned* flag_limit_out, cudaStream_t st) {
  const TcLayout l = tc_layout(N, P1, P2);
  char* base = reinterpret_cast<char*>(ws);
  unsigned* maxw = reinterpret_cast<unsigned*>(base + l.maxw_off);
  float* w = reinterpret_cast<float*>(base + l.w_off);
  unsigned* maxq = reinterpret_cast<unsigned*>(base + l.maxq_off);
  float* xq = reinterpret_cast<float*>(base + l.xq_off);
  uint2* cands = reinterpret_cast<uint2*>(base + l.cands_off);
  unsigned* counts = reinterpret_cast<unsigned*>(base + l.counts_off);
  float* tfin = reinterpret_cast<float*>(base + l.tfin_off);
  unsigned char* flags = reinterpret_cast<unsigned char*>(base + l.flags_off);
  *flags_out = flags;
  unsigned* flag_rows = reinterpret_cast<unsigned*>(base + l.rows_off + 256);
  *flag_count_out = flag_rows - 1;
  const unsigned limit = 4096;  // flagged rows the per-row kernel takes; more -> dense generic kernel
  *flag_limit_out = limit;
  POPS_CUDA_OK(cudaMemsetAsync(flag_rows - 1, 0, 4, st));

  POPS_CUDA_OK(cudaMemsetAsync(maxw, 0, size_t(N) * 4, st));
  POPS_CUDA_OK(cudaMemsetAsync(maxq, 0, size_t(N) * 4, st));
  {
    dim3 grid(static_cast<unsigned>(ceil_div(l.P2pad, 8)), N);
    tc_norm_kernel<<<grid, 256, 0, st>>>(p2, len2, P2, l.P2pad, D, w, maxw);
    POPS_LAUNCH_OK("tc_norm_kernel");
    dim3 gridq(static_cast<unsigned>(ceil_div(l.P1pad, 8)), N);
    tc_norm_kernel<<<gridq, 256, 0, st>>>(p1, len1, P1, l.P1pad, D, xq, maxq);
    POPS_LAUNCH_OK("tc_norm_kernel");
  }
  const int cl_env = get_option("tc_cluster", 1);  // CTAs sharing every p2 stage by TMA multicast (tuning aid)
  const int64_t qtiles = ceil_div(P1, TC_M);
  const bool pair = get_option("tc_pair", 1) != 0 && qtiles >= 2;  // cta_group::2 pairs
  const int CL = pair ? 1 : ((cl_env >= 4 && qtiles >= 4) ? 4 : ((cl_env >= 2 && qtiles >= 2) ? 2 : 1));
  CUtensorMap map_q, map_p;
  int rc = make_map(&map_q, p1, N, P1, D, TC_M);
  if (rc != POPS_OK) return rc;
  rc = make_map(&map_p, p2, N, P2, D, TC_N / CL);
  if (rc != POPS_OK) return rc;

  TcParams prm;
  prm.w = w; prm.len1 = len1; prm.len2 = len2; prm.xq = xq; prm.maxw_bits = maxw; prm.cands = cands; prm.counts = counts; prm.tfin = tfin;
  prm.P1 = P1; prm.P2 = P2; prm.P2pad = l.P2pad; prm.P1pad = l.P1pad; prm.D = D;
  prm.KB = (D + TC_KBLK - 1) / TC_KBLK;
  prm.dbg = get_option("tc_dbg", 0);
  prm.seed_tiles = std::max(1, std::min(64, get_option("tc_seed", 4)));
  const size_t fixed = tc_smem_fixed(prm.KB);
  prm.nstage = static_cast<int>(std::min<size_t>(TC_MAX_STAGES, (kSmemLimit - fixed) / TC_STAGE_BYTES));
  const size_t smem = fixed + size_t(prm.nstage) * TC_STAGE_BYTES;
  {
    void (*kern)(const CUtensorMap, const CUtensorMap, const TcParams) =
        pair ? knn_tc_scan_kernel<1, true>
             : (CL == 4 ? knn_tc_scan_kernel<4, false>
                        : (CL == 2 ? knn_tc_scan_kernel<2, false> : knn_tc_scan_kernel<1, false>));
    const int cdim = pair ? 2 : CL;
    POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(ceil_div(qtiles, cdim) * cdim), N);  // whole clusters; spare CTAs see no query
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cdim;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    profile_begin("knn_tc_scan", st);
    POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_q, map_p, prm));
    profile_end("knn_tc_scan", st);
    POPS_LAUNCH_OK("knn_tc_scan_kernel");
  }
  TcRerankParams rp;
  rp.p1 = p1; rp.p2 = p2; rp.len1 = len1; rp.len2 = len2; rp.cands = cands; rp.counts = counts; rp.tfin = tfin; rp.xq = xq; rp.maxw_bits = maxw; rp.maxq_bits = maxq;
  rp.idx = idx; rp.dists = dists; rp.flags = flags; rp.flag_rows = flag_rows; rp.P1 = P1; rp.P2 = P2; rp.P1pad = l.P1pad; rp.D = D; rp.K = K;
  rp.debug = get_option("knn_stats", 0);
  {
    dim3 grid(static_cast<unsigned>(ceil_div(P1, 4)), N);
    profile_begin("knn_tc_rerank", st);
    knn_tc_rerank_kernel<<<grid, 128, size_t(4) * D * 4, st>>>(rp);
    profile_end("knn_tc_rerank", st);
    POPS_LAUNCH_OK("knn_tc_rerank_kernel");
  }
  {
    const size_t esmem = align_up(size_t(D) * 4, 16) + 50 * 8 + size_t(kExactThreads / 32) * 32 * 33 * 4;
    POPS_CUDA_OK(cudaFuncSetAttribute(knn_exact_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(esmem)));
    POPS_CUDA_OK(cudaFuncSetAttribute(knn_exact_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(esmem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(std::max(1, num_sms() / kExactCluster) * kExactCluster));
    cfg.blockDim = dim3(kExactThreads);
    cfg.dynamicSmemBytes = esmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kExactCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const unsigned* frows = flag_rows;
    profile_begin("knn_exact_rows", st);
    if (K <= 4)
      POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, knn_exact_rows_kernel<4>, p1, p2, len2, frows, limit, P1, P2, D, K, idx, dists));
    else
      POPS_CUDA_OK(cudaLaunchKernelEx(&cfg, knn_exact_rows_kernel<16>, p1, p2, len2, frows, limit, P1, P2, D, K, idx, dists));
    profile_end("knn_exact_rows", st);
    POPS_LAUNCH_OK("knn_exact_rows_kernel");
  }
  return POPS_OK;
}

}  // namespace pops
