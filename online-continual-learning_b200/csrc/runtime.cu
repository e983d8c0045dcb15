// runtime.cu -- error string, launch counter, device properties.
#include <stdarg.h>

#include "common.cuh"

namespace b200ocl {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200ocl

extern "C" {
const char* b200ocl_last_error(void) { return b200ocl::g_err; }
int b200ocl_version(void) { return 100; }
uint64_t b200ocl_launch_count(void) { return b200ocl::g_launches.load(std::memory_order_relaxed); }
}
