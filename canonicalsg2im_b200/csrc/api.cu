// Error channel, version and device queries of the csg2im C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void csg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int csg_num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

static unsigned long long g_launches = 0;
void csg_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// number of kernel-launch sites passed since load (a launch site may issue 1-2 kernels); bench.py reports it
CSG_API long long csg_launch_count(void) { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

CSG_API const char* csg_last_error(void) { return g_err; }
CSG_API void csg_clear_error(void) { g_err[0] = 0; }
CSG_API int csg_version(void) { return 100; }   // 0.1.0
CSG_API int csg_device_sms(void) { return csg_num_sms(); }
