// Probe of cp.async.bulk.tensor.2d.tile::gather4 semantics on sm_100a (box rows, swizzle, coordinate order).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3, uint16_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* bp = smem + (base - smem_u32(smem));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) bp[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(base), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar))
        : "memory");
    uint32_t ok = 0; int spins = 0;
    while (!ok && spins < (1 << 22)) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      ++spins;
    }
    *status = ok ? 1 : -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(bp)[i];
}
int main(int argc, char** argv) {
  int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  int swz = argc > 2 ? atoi(argv[2]) : 1;
  const int R = 500, C = 128;
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 64 + (c & 63) + ((c >> 6) << 15));  // unique-ish code
  uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[2] = {C, R}; cuuint64_t strides[1] = {C * 2}; cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d swizzle=%d -> %d\n", box_rows, swz, (int)r);
  if (r) return 0;
  uint16_t* out; int* st; cudaMalloc(&out, 2048); cudaMalloc(&st, 4); cudaMemset(st, 0, 4);
  int rows[4] = {5, 400, 17, 3}, col = 64;
  probe<<<1, 128, 4096>>>(tm, col, rows[0], rows[1], rows[2], rows[3], out, st);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e) return 0;
  uint16_t ho[1024]; int hs; cudaMemcpy(ho, out, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
  printf("status %d\n", hs);
  // for each dst row i (128 B = 64 elements) and 16B chunk j, report which (row, col-chunk) of the source it holds
  for (int i = 0; i < 8; ++i) {
    printf("dst row %d:", i);
    for (int j = 0; j < 8; ++j) {
      uint16_t v = ho[i * 64 + j * 8];
      if (v == 0xEEEE) { printf("  ----"); continue; }
      int hi = v >> 15, rr = (v & 0x7FFF) / 64, cc = (v & 63);
      printf("  r%d.c%d", rr, hi * 64 + cc);
    }
    printf("\n");
  }
  return 0;
}
