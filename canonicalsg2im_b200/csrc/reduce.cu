// One launch for the final passes of several deterministic reductions (see internal.h).
#include "internal.h"

namespace {

struct ReduceJobs {
  CsgReduceJob job[CSG_REDUCE_MAX_JOBS];
  int block_end[CSG_REDUCE_MAX_JOBS];   // exclusive prefix of the blocks of each job
  int n;
};

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
  if (q.lanes == 1) {
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
    blocks += csg_div_up(q.n, q.lanes == 1 ? 256 : 32);
    r.job[r.n] = q;
    r.block_end[r.n] = blocks;
    ++r.n;
  }
  if (r.n == 0) return 0;
  for (int k = r.n; k < CSG_REDUCE_MAX_JOBS; ++k) { r.job[k] = r.job[0]; r.block_end[k] = blocks; }
  CSG_CUDA(csg_launch_pdl(reduce_multi_kernel, dim3(blocks), dim3(256), 0, stream, r));
  CSG_CHECK_LAUNCH("csg_reduce_multi");
  return 0;
}
