// Library-level plumbing of libpointops_b200.so: error string, build info, launch counter.
#include "common.cuh"

namespace pops {
std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}
std::atomic<int64_t> g_launch_count{0};
}  // namespace pops

extern "C" int pops_abi_version(void) { return 1; }

extern "C" const char* pops_build_info(void) {
  return "libpointops_b200 abi=1 arch=sm_100a cuda="
#define POPS_STR2(x) #x
#define POPS_STR(x) POPS_STR2(x)
      POPS_STR(__CUDACC_VER_MAJOR__) "." POPS_STR(__CUDACC_VER_MINOR__) " built " __DATE__;
}

extern "C" const char* pops_last_error(void) { return pops::last_error_ref().c_str(); }

extern "C" int64_t pops_launch_count(void) {
  return pops::g_launch_count.load(std::memory_order_relaxed);
}
