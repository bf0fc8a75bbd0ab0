// Multi-tensor Adam update for the parameters of the scene-graph -> layout model (the reference trains them with
// torch.optim.Adam, scripts/train.py): one launch updates up to 48 fp32 tensors, every element read and written once
// with 128-bit accesses.  Same arithmetic as torch.optim.Adam (amsgrad = False, maximize = False):
//   g' = g + wd * p;  m = m + (g' - m) * (1 - b1);  v = b2 * v + (1 - b2) * g'^2
//   p  = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace {

constexpr int ADAM_MAX = 48;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = ADAM_THREADS * 4 * 4;     // elements per block: 4 float4 per thread

struct AdamJobs {
  float* p[ADAM_MAX];
  const float* g[ADAM_MAX];
  float* m[ADAM_MAX];
  float* v[ADAM_MAX];
  int n[ADAM_MAX];
  float lr_over_bc1[ADAM_MAX];     // per tensor: torch.optim.Adam keeps one step count per parameter
  float inv_bc2_sqrt[ADAM_MAX];
};

struct AdamScalars {
  float lr_over_bc1, inv_bc2_sqrt, b1, b2, eps, wd;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s) {
  g = fmaf(s.wd, p, g);
  m = m + (g - m) * (1.f - s.b1);
  v = s.b2 * v + (1.f - s.b2) * g * g;
  const float denom = sqrtf(v) * s.inv_bc2_sqrt + s.eps;
  p = p - s.lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(AdamJobs jobs, AdamScalars s) {
  CSG_PDL_WAIT();
  const int t = blockIdx.y;
  s.lr_over_bc1 = jobs.lr_over_bc1[t];
  s.inv_bc2_sqrt = jobs.inv_bc2_sqrt[t];
  const int n = jobs.n[t];
  const int beg = blockIdx.x * ADAM_CHUNK;
  if (beg >= n) return;
  float* __restrict__ p = jobs.p[t];
  const float* __restrict__ g = jobs.g[t];
  float* __restrict__ m = jobs.m[t];
  float* __restrict__ v = jobs.v[t];
  const int end = min(n, beg + ADAM_CHUNK);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (aligned) {
    const int end4 = beg + ((end - beg) & ~3);
    for (int i = beg + threadIdx.x * 4; i < end4; i += ADAM_THREADS * 4) {
      float4 pp = ld_f4(p + i), gg = ld_f4(g + i), mm = ld_f4(m + i), vv = ld_f4(v + i);
      adam_one(pp.x, gg.x, mm.x, vv.x, s);
      adam_one(pp.y, gg.y, mm.y, vv.y, s);
      adam_one(pp.z, gg.z, mm.z, vv.z, s);
      adam_one(pp.w, gg.w, mm.w, vv.w, s);
      st_f4(p + i, pp); st_f4(m + i, mm); st_f4(v + i, vv);
    }
    for (int i = end4 + threadIdx.x; i < end; i += ADAM_THREADS) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_one(pp, g[i], mm, vv, s);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  } else {
    for (int i = beg + threadIdx.x; i < end; i += ADAM_THREADS) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_one(pp, g[i], mm, vv, s);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

}  // namespace

// count tensors (HOST arrays of device pointers / element counts / per-tensor update indices t >= 1).
CSG_API int csg_adam_multi(int count, void* const* params, const void* const* grads, void* const* exp_avg,
                           void* const* exp_avg_sq, const int* numel, double lr, double beta1, double beta2, double eps,
                           double weight_decay, const int* steps, cudaStream_t stream) {
  CSG_REQUIRE(count >= 0 && steps, "adam_multi: bad count=%d / steps", count);
  AdamScalars s;
  s.lr_over_bc1 = 0.f; s.inv_bc2_sqrt = 0.f;
  s.b1 = (float)beta1; s.b2 = (float)beta2; s.eps = (float)eps; s.wd = (float)weight_decay;
  for (int base = 0; base < count; base += ADAM_MAX) {
    const int k = count - base < ADAM_MAX ? count - base : ADAM_MAX;
    AdamJobs jobs;
    int max_n = 0;
    for (int i = 0; i < ADAM_MAX; ++i) {
      const bool on = i < k;
      jobs.p[i] = on ? reinterpret_cast<float*>(params[base + i]) : nullptr;
      jobs.g[i] = on ? reinterpret_cast<const float*>(grads[base + i]) : nullptr;
      jobs.m[i] = on ? reinterpret_cast<float*>(exp_avg[base + i]) : nullptr;
      jobs.v[i] = on ? reinterpret_cast<float*>(exp_avg_sq[base + i]) : nullptr;
      jobs.n[i] = on ? numel[base + i] : 0;
      jobs.lr_over_bc1[i] = 0.f; jobs.inv_bc2_sqrt[i] = 0.f;
      if (on) {
        const int step = steps[base + i];
        CSG_REQUIRE(step >= 1, "adam_multi: tensor %d has step %d < 1", base + i, step);
        const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
        jobs.lr_over_bc1[i] = (float)(lr / bc1);
        jobs.inv_bc2_sqrt[i] = (float)(1.0 / sqrt(bc2));
        CSG_REQUIRE(numel[base + i] >= 0 && (numel[base + i] == 0 || (jobs.p[i] && jobs.g[i] && jobs.m[i] && jobs.v[i])),
                    "adam_multi: tensor %d has a null pointer", base + i);
        if (jobs.n[i] > max_n) max_n = jobs.n[i];
      }
    }
    if (max_n == 0) continue;
    CSG_CUDA(csg_launch_pdl(adam_multi_kernel, dim3(dim3(csg_div_up(max_n, ADAM_CHUNK), k)), dim3(ADAM_THREADS), 0, stream, jobs, s));
    CSG_CHECK_LAUNCH("csg_adam_multi");
  }
  return 0;
}
