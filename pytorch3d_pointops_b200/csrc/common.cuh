// Shared device/host helpers for libpointops_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/pointops_b200.h"

namespace pops {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
std::string& last_error_ref();
extern std::atomic<int64_t> g_launch_count;

inline int fail(int code, const std::string& msg) {
  last_error_ref() = msg;
  return code;
}

#define POPS_CHECK_ARG(cond, msg)                                                     \
  do {                                                                                \
    if (!(cond)) return ::pops::fail(POPS_ERR_INVALID_ARGUMENT, std::string(msg));    \
  } while (0)

#define POPS_CUDA_OK(expr)                                                            \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ::pops::fail(POPS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// call after every kernel launch (mirrors AT_CUDA_CHECK(cudaGetLastError()), knn.cu:457)
#define POPS_LAUNCH_OK(name)                                                          \
  do {                                                                                \
    ::pops::g_launch_count.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess)                                                            \
      return ::pops::fail(POPS_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// tuning / measurement knobs (api.cu): value set through pops_set_option, else the environment
// variable POPS_<NAME>, else `dflt`.  Read at every call, so tests and bench.py can flip them.
int get_option(const char* name, int dflt);

// per-kernel timing (api.cu); no-ops unless pops_profile_enable(1)
void profile_begin(const char* kernel, cudaStream_t st);
void profile_end(const char* kernel, cudaStream_t st);

inline int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

#ifdef __CUDACC__
#define POPS_HD __host__ __device__
#else
#define POPS_HD
#endif
POPS_HD inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
POPS_HD inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// same, for single-thread roles that share a scheduler with working warps: the suspend-time hint
// lets the hardware park the thread until the phase completes (or the hint expires) instead of
// burning issue slots in a polling loop
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "PARK_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra PARK_DONE;\n\t"
      "bra PARK_LOOP;\n\t"
      "PARK_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// 1-D bulk global->shared copy completing on an mbarrier.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// order prior generic-proxy smem accesses before subsequent async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- exact (never contracted) distance arithmetic -----------------------------------------
// The reference's CPU build has no FMA: d = fl(fl(dx*dx) + fl(dy*dy)) + ...
template <int NORM>
__device__ __forceinline__ float dist_term(float a, float b) {
  const float diff = __fsub_rn(a, b);
  return NORM == 2 ? __fmul_rn(diff, diff) : fabsf(diff);
}

__device__ __forceinline__ uint64_t make_key(float d, uint32_t j) {
  return (static_cast<uint64_t>(__float_as_uint(d)) << 32) | j;
}
__device__ __forceinline__ float key_dist(uint64_t k) {
  return __uint_as_float(static_cast<uint32_t>(k >> 32));
}

// ---- non-finite / out-of-range inputs ------------------------------------------------------------
// The filtered searches (expanded-form filter, box pruning, tensor-core filter) are proven for finite
// coordinates with |c| < 1e18.  Every pre-pass therefore records, per cloud, the
// largest |coordinate| as an unsigned BIT PATTERN (integer max: +inf and every NaN compare above all
// finite values, unlike fmaxf, which drops NaN); a cloud at or above kDirtyBits is skipped by the fast
// kernels and answered by the exact generic kernel, whose keys order  finite < +inf < NaN, ties by
// lower index.  Where the reference's result is well defined (NaN query: the first K points; +inf
// distances: ordinary values) this is the reference's result (knn_cpu.cpp:40-65); a NaN POINT makes
// the reference's heap comparator inconsistent -- there NaN simply ranks last here.
constexpr unsigned kDirtyBits = 0x5d5e0b6bu;      // 1e18f
constexpr unsigned kDirtyNormBits = 0x7b4097ceu;  // 1e36f: squared norms (tensor-core path)
__device__ __forceinline__ unsigned abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
// key with the total order above: NaN distances are canonicalised so that they tie among themselves
__device__ __forceinline__ uint64_t make_key_total(float d, uint32_t j) {
  const uint32_t b = (d != d) ? 0x7fffffffu : __float_as_uint(d);
  return (static_cast<uint64_t>(b) << 32) | j;
}

// One 12-byte row (x, y, z) with two loads instead of three: an 8-byte and a 4-byte one, whichever order
// the row's alignment allows.  Row gathers are bound by L1 sector lookups per instruction, not by bytes.
__device__ __forceinline__ void ldg_row3(const float* row, float& x, float& y, float& z) {
  const bool al = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
  const float2 a = __ldg(reinterpret_cast<const float2*>(al ? row : row + 1));
  const float b = __ldg(al ? row + 2 : row);
  x = al ? a.x : b;
  y = al ? a.y : a.x;
  z = al ? b : a.y;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace pops
