// Error channel, version and device queries of the csg2im C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <mutex>

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

// ------------------------------------------------------------------------------------------------
// Asynchronous index errors.  Kernels that use caller-supplied integers as offsets (triple subject / object /
// predicate ids, embedding ids, canonicalization triplets) cannot raise; they neutralise the offending row and
// report it through one pinned, device-mapped host record {code, row, value, limit}.  The host reads the record
// without synchronising (csg_async_error_poll): the reference raises IndexError at the same inputs, here the error
// surfaces at the first library call after the kernel has run (like a CUDA device-side assert, but recoverable).
// ------------------------------------------------------------------------------------------------
static int* g_async_rec = nullptr;
int* csg_async_err_ptr() {
  static std::once_flag once;
  std::call_once(once, [] {
    int* p = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), 8 * sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
      for (int i = 0; i < 8; ++i) p[i] = 0;
      g_async_rec = p;
    } else {
      cudaGetLastError();
    }
  });
  return g_async_rec;   // unified addressing: the host pointer is valid on every device
}
// out4 (HOST, may be NULL) <- {code, row, value, limit}; returns code (0 = none) and clears the record.
CSG_API int csg_async_error_poll(int* out4) {
  int* p = csg_async_err_ptr();
  if (!p) return 0;
  volatile int* v = p;
  const int code = v[0];
  if (out4) for (int i = 0; i < 4; ++i) out4[i] = v[i];
  if (code) for (int i = 0; i < 4; ++i) v[i] = 0;
  return code;
}

CSG_API const char* csg_last_error(void) { return g_err; }
CSG_API void csg_clear_error(void) { g_err[0] = 0; }
CSG_API int csg_version(void) { return 100; }   // 0.1.0
CSG_API int csg_device_sms(void) { return csg_num_sms(); }

// ------------------------------------------------------------------------------------------------
// live timing of kernel classes (bench.py's roofline legs).  Off by default; when on, every instrumented
// entry point records two CUDA events on ITS stream; csg_prof_collect synchronises them and sums per class.
// ------------------------------------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace {
struct ProfRec { cudaEvent_t e0, e1; int cls; double work; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<ProfRec> g_prof_pool;
int g_prof_on = 0;
}  // namespace

CsgProfScope::CsgProfScope(int cls, double work, cudaStream_t s) : slot(-1), stream(s) {
  if (!__atomic_load_n(&g_prof_on, __ATOMIC_RELAXED)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  if (!g_prof_pool.empty()) { r = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  r.cls = cls; r.work = work;
  cudaEventRecord(r.e0, s);
  slot = (int)g_prof_recs.size();
  g_prof_recs.push_back(r);
}
CsgProfScope::~CsgProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot].e1, stream);
}

CSG_API int csg_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) g_prof_pool.push_back(r);
  g_prof_recs.clear();
  __atomic_store_n(&g_prof_on, on ? 1 : 0, __ATOMIC_RELAXED);
  return 0;
}
// out: HOST array [CSG_PROF_CLASSES][3] = {work, seconds, scopes} per class; clears the records.
CSG_API int csg_prof_collect(double* out) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < CSG_PROF_CLASSES * 3; ++i) out[i] = 0.0;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      out[r.cls * 3 + 0] += r.work;
      out[r.cls * 3 + 1] += (double)ms * 1e-3;
      out[r.cls * 3 + 2] += 1.0;
    }
    g_prof_pool.push_back(r);
  }
  g_prof_recs.clear();
  return 0;
}
