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

CSG_API const char* csg_last_error(void) { return g_err; }
CSG_API void csg_clear_error(void) { g_err[0] = 0; }
CSG_API int csg_version(void) { return 100; }   // 0.1.0
CSG_API int csg_device_sms(void) { return csg_num_sms(); }
