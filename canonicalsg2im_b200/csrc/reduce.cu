// One launch for the final passes of several deterministic reductions (see internal.h).
#include "internal.h"

namespace {

struct ReduceJobs {
  CsgReduceJob job[CSG_REDUCE_MAX_JOBS];
  int block_end[CSG_REDUCE_MAX_JOBS];   // exclusive prefix of the blocks of each job
  int vec[CSG_REDUCE_MAX_JOBS];         // lanes == 1 jobs: four consecutive outputs per thread through 16-byte loads
  int n;
};

// four independent ordered sums side by side (same order per output as ordered_sum): 16-byte loads, U in flight
template <int U>
__device__ __forceinline__ float4 ordered_sum4(const float* __restrict__ p, size_t stride, int n) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int b = 0;
  for (; b + U <= n; b += U) {
    float4 x[U];
#pragma unroll
    for (int i = 0; i < U; ++i) x[i] = ld_f4(p + (size_t)(b + i) * stride);
#pragma unroll
    for (int i = 0; i < U; ++i) { s.x += x[i].x; s.y += x[i].y; s.z += x[i].z; s.w += x[i].w; }
  }
  for (; b < n; ++b) {
    const float4 x = ld_f4(p + (size_t)b * stride);
    s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
  }
  return s;
}

__global__ void __launch_bounds__(256) reduce_multi_kernel(const ReduceJobs jobs) {
  CSG_PDL_WAIT();
  __shared__ float red[8][33];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.block_end[j]) ++j;
  const CsgReduceJob& q = jobs.job[j];
  const int blk = (int)blockIdx.x - (j ? jobs.block_end[j - 1] : 0);
  float s = 0.f;
  int i;
  bool write;
  if (q.lanes == 1 && jobs.vec[j]) {
    i = (blk * 256 + threadIdx.x) * 4;
    if (i < q.n) {
      const float4 v = ordered_sum4<8>(q.partial + i, (size_t)q.stride, q.parts);
      st_f4(q.out + (q.ncols > 0 ? (size_t)(i / q.ncols) * q.ldo + (i % q.ncols) : (size_t)i), v);
    }
    return;
  } else if (q.lanes == 1) {
    i = blk * 256 + threadIdx.x;
    write = i < q.n;
    if (write) s = ordered_sum<8>(q.partial + i, (size_t)q.stride, q.parts);
  } else {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    i = blk * 32 + tx;
    float acc = 0.f;
    if (i < q.n && ty < q.parts) acc = ordered_sum<8>(q.partial + (size_t)ty * q.stride + i, (size_t)8 * q.stride, (q.parts - ty + 7) / 8);
    red[ty][tx] = acc;
    __syncthreads();
    write = ty == 0 && i < q.n;
    if (write) {
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][tx];
    }
  }
  if (write) {
    if (q.op == CSG_RED_SIGMOID_GRAD) {
      const float sg = 1.f / (1.f + expf(-q.aux[i]));
      s = s * sg * (1.f - sg);
    }
    q.out[q.ncols > 0 ? (size_t)(i / q.ncols) * q.ldo + (i % q.ncols) : (size_t)i] = s;
  }
}

}  // namespace

int csg_reduce_multi(const CsgReduceJob* jobs, int njobs, cudaStream_t stream) {
  CSG_REQUIRE(njobs >= 0 && njobs <= CSG_REDUCE_MAX_JOBS, "reduce_multi: %d jobs", njobs);
  ReduceJobs r;
  r.n = 0;
  int blocks = 0;
  for (int k = 0; k < njobs; ++k) {
    const CsgReduceJob& q = jobs[k];
    if (q.parts <= 0 || q.n <= 0) continue;
    CSG_REQUIRE(q.partial && q.out && (q.lanes == 1 || q.lanes == 8), "reduce_multi: bad job %d", k);
    // plain sums whose rows, pitches and pointers keep 16-byte alignment take four outputs per thread
    const bool vec = q.lanes == 1 && q.op == CSG_RED_SUM && (q.n & 3) == 0 && (q.stride & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(q.partial) | reinterpret_cast<uintptr_t>(q.out)) & 15) == 0 &&
                     (q.ncols == 0 || ((q.ncols & 3) == 0 && (q.ldo & 3) == 0));
    blocks += csg_div_up(q.n, q.lanes == 1 ? (vec ? 1024 : 256) : 32);
    r.vec[r.n] = vec ? 1 : 0;
    r.job[r.n] = q;
    r.block_end[r.n] = blocks;
    ++r.n;
  }
  if (r.n == 0) return 0;
  for (int k = r.n; k < CSG_REDUCE_MAX_JOBS; ++k) { r.job[k] = r.job[0]; r.block_end[k] = blocks; r.vec[k] = 0; }
  CSG_CUDA(csg_launch_pdl(reduce_multi_kernel, dim3(blocks), dim3(256), 0, stream, r));
  CSG_CHECK_LAUNCH("csg_reduce_multi");
  return 0;
}
