// Library-level plumbing of the C ABI: error text, launch accounting, device queries.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace pp

extern "C" int pp_abi_version(void) { return PP_ABI_VERSION; }

extern "C" const char* pp_last_error(void) { return pp::g_err; }

#define PP_STR2(x) #x
#define PP_STR(x) PP_STR2(x)
extern "C" const char* pp_build_info(void) {
  return "poseprobe_b200 abi " PP_STR(PP_ABI_VERSION) " sm_100a nvcc " PP_STR(__CUDACC_VER_MAJOR__) "." PP_STR(
      __CUDACC_VER_MINOR__) "." PP_STR(__CUDACC_VER_BUILD__) " built " __DATE__;
}

extern "C" int pp_check_device_ptr(const void* p) {
  if (!p) {
    pp::set_error("null pointer");
    return PP_ERR_INVALID_ARG;
  }
  cudaPointerAttributes a;
  memset(&a, 0, sizeof(a));
  const cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    pp::set_error("cudaPointerGetAttributes: %s", cudaGetErrorString(e));
    return PP_ERR_NOT_DEVICE;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return PP_OK;
  pp::set_error("pointer %p is not device memory (no CPU fallback)", p);
  return PP_ERR_NOT_DEVICE;
}

extern "C" uint64_t pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }
