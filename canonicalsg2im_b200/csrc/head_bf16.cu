// Narrow output head of box_net (sg2im/model.py:58-60: Linear(D, H) -> ReLU -> Linear(H, 4)) for the bf16 engine.
// The wide first layer runs on the tcgen05 GEMMs; the 4-wide second layer is far below any GEMM tile, so it is a
// row-dot kernel forward and one fused kernel backward:
//   csg_head_fwd   y[r, j] = sum_k h[r, k] * w[j, k] + b[j]                       (h bf16, w / b / y fp32, j < NOUT <= 8)
//   csg_head_bwd   dh[r, k] = (h[r, k] > 0) * sum_j dy[r, j] * w[j, k]  (bf16, ReLU of the first layer folded in),
//                  dw[j, k] = sum_r dy[r, j] * h[r, k],  db[j] = sum_r dy[r, j]  (per-block partials summed in block
//                  order by a second kernel: deterministic)
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int HEAD_MAX_OUT = 8;
constexpr int HB_ROWS = 32;            // rows per block of the backward kernel

__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ h, int ldh, const float* __restrict__ w,
                                                       const float* __restrict__ b, int M, int K, int NOUT,
                                                       float* __restrict__ y, bool h_fp16) {
  CSG_PDL_WAIT();
  extern __shared__ __align__(16) float ws[];          // [NOUT][K]
  for (int i = threadIdx.x; i < NOUT * K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < M; r += gridDim.x * 8) {
    float acc[HEAD_MAX_OUT];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_OUT; ++j) acc[j] = 0.f;
    for (int k = lane * 2; k < K; k += 64) {
      const float2 x = h_fp16 ? __half22float2(*reinterpret_cast<const __half2*>(h + (size_t)r * ldh + k))
                              : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + (size_t)r * ldh + k));
#pragma unroll
      for (int j = 0; j < HEAD_MAX_OUT; ++j)
        if (j < NOUT) acc[j] = fmaf(x.y, ws[j * K + k + 1], fmaf(x.x, ws[j * K + k], acc[j]));
    }
#pragma unroll
    for (int j = 0; j < HEAD_MAX_OUT; ++j)
      if (j < NOUT) {
        const float v = warp_sum(acc[j]);
        if (lane == 0) y[(size_t)r * NOUT + j] = v + b[j];
      }
  }
}

// thread = column k; a block walks HB_ROWS rows
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ h, int ldh,
                                                       const float* __restrict__ w, int M, int K, int NOUT,
                                                       __nv_bfloat16* __restrict__ dh, int lddh, float* __restrict__ partial,
                                                       bool h_fp16) {
  CSG_PDL_WAIT();
  __shared__ float sdy[HB_ROWS][HEAD_MAX_OUT];
  const int r0 = blockIdx.x * HB_ROWS, nr = min(HB_ROWS, M - r0);
  for (int i = threadIdx.x; i < HB_ROWS * HEAD_MAX_OUT; i += blockDim.x) {
    const int r = i / HEAD_MAX_OUT, j = i % HEAD_MAX_OUT;
    sdy[r][j] = (r < nr && j < NOUT) ? dy[(size_t)(r0 + r) * NOUT + j] : 0.f;
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (NOUT * K + HEAD_MAX_OUT);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float wk[HEAD_MAX_OUT], acc[HEAD_MAX_OUT];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_OUT; ++j) {
      wk[j] = j < NOUT ? w[(size_t)j * K + k] : 0.f;
      acc[j] = 0.f;
    }
    for (int r = 0; r < nr; ++r) {
      const float hv = h_fp16 ? __half2float(reinterpret_cast<const __half*>(h)[(size_t)(r0 + r) * ldh + k])
                              : __bfloat162float(h[(size_t)(r0 + r) * ldh + k]);
      float g = 0.f;
#pragma unroll
      for (int j = 0; j < HEAD_MAX_OUT; ++j) {
        g = fmaf(sdy[r][j], wk[j], g);
        acc[j] = fmaf(sdy[r][j], hv, acc[j]);
      }
      dh[(size_t)(r0 + r) * lddh + k] = __float2bfloat16_rn(hv > 0.f ? g : 0.f);
    }
#pragma unroll
    for (int j = 0; j < HEAD_MAX_OUT; ++j)
      if (j < NOUT) out[(size_t)j * K + k] = acc[j];
  }
  if (threadIdx.x < HEAD_MAX_OUT) {
    float sdb = 0.f;
    for (int r = 0; r < nr; ++r) sdb += sdy[r][threadIdx.x];
    out[(size_t)NOUT * K + threadIdx.x] = sdb;
  }
}

__global__ void head_bwd_final_kernel(const float* __restrict__ partial, int blocks, int K, int NOUT, float* __restrict__ dw,
                                      float* __restrict__ db) {
  CSG_PDL_WAIT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cells = NOUT * K + HEAD_MAX_OUT;
  if (i >= cells) return;
  const float a = ordered_sum<16>(partial + i, (size_t)cells, blocks);
  if (i < NOUT * K) dw[i] = a;
  else if (i - NOUT * K < NOUT) db[i - NOUT * K] = a;
}

}  // namespace

CSG_API int csg_head_fwd(const void* h, int ldh, const float* w, const float* b, int M, int K, int nout, float* y,
                         int h_fp16, cudaStream_t stream) {
  if (M == 0) return 0;
  CSG_REQUIRE(nout >= 1 && nout <= HEAD_MAX_OUT && K > 0 && (K & 1) == 0 && (ldh & 1) == 0,
              "head_fwd: nout=%d must be in [1, 8], K=%d and ldh=%d even", nout, K, ldh);
  const size_t smem = (size_t)nout * K * sizeof(float);
  CSG_REQUIRE(smem <= 48 * 1024, "head_fwd: weight [%d, %d] does not fit shared memory", nout, K);
  int blocks = csg_div_up(M, 8);
  if (blocks > 4 * csg_num_sms()) blocks = 4 * csg_num_sms();
  CSG_CUDA(csg_launch_pdl(head_fwd_kernel, dim3(blocks), dim3(256), smem, stream, reinterpret_cast<const __nv_bfloat16*>(h), ldh, w, b, M, K, nout, y, h_fp16 != 0));
  CSG_CHECK_LAUNCH("csg_head_fwd");
  return 0;
}

CSG_API size_t csg_head_bwd_workspace(int M, int K, int nout) {
  return (size_t)csg_div_up(M > 0 ? M : 1, HB_ROWS) * ((size_t)nout * K + HEAD_MAX_OUT) * sizeof(float) + 256;
}

// dh [M, K] bf16 (pitch lddh), dw [nout, K] fp32, db [nout] fp32 are fully written.
CSG_API int csg_head_bwd(const float* dy, const void* h, int ldh, const float* w, int M, int K, int nout, void* dh, int lddh,
                         float* dw, float* db, int h_fp16, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  CSG_REQUIRE(nout >= 1 && nout <= HEAD_MAX_OUT && K > 0, "head_bwd: nout=%d must be in [1, 8]", nout);
  if (M == 0) {
    CSG_CUDA(cudaMemsetAsync(dw, 0, (size_t)nout * K * sizeof(float), stream));
    CSG_CUDA(cudaMemsetAsync(db, 0, (size_t)nout * sizeof(float), stream));
    return 0;
  }
  CSG_REQUIRE(workspace && workspace_bytes >= csg_head_bwd_workspace(M, K, nout), "head_bwd: workspace too small");
  const int blocks = csg_div_up(M, HB_ROWS);
  float* partial = reinterpret_cast<float*>(workspace);
  CSG_CUDA(csg_launch_pdl(head_bwd_kernel, dim3(blocks), dim3(256), 0, stream, dy, reinterpret_cast<const __nv_bfloat16*>(h), ldh, w, M, K, nout,
                                              reinterpret_cast<__nv_bfloat16*>(dh), lddh, partial, h_fp16 != 0));
  CSG_CHECK_LAUNCH("csg_head_bwd");
  const int cells = nout * K + HEAD_MAX_OUT;
  CSG_CUDA(csg_launch_pdl(head_bwd_final_kernel, dim3(csg_div_up(cells, 256)), dim3(256), 0, stream, partial, blocks, K, nout, dw, db));
  CSG_CHECK_LAUNCH("csg_head_bwd final");
  return 0;
}
