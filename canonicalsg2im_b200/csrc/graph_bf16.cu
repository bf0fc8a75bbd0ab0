// bf16-activation variants of the pooling / assembly kernels used by the tensor-core engine, plus
// the fp32 -> bf16 weight cast (with the transposed copy the dX-type GEMMs consume).
// Semantics are those of graph.cu (sg2im/graph.py:69-107); accumulation stays fp32.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack2(f[0], f[1]); u.y = pack2(f[2], f[3]); u.z = pack2(f[4], f[5]); u.w = pack2(f[6], f[7]);
  return u;
}

// dst[r, c] (or dst[c, r] when transpose) = bf16(src[r, c])
__global__ void cast_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld_src,
                                 __nv_bfloat16* __restrict__ dst, int ld_dst, int transpose) {
  __shared__ float tile[32][33];
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    if (!transpose) {
      int r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) dst[(size_t)r * ld_dst + c] = __float2bfloat16_rn(tile[i][threadIdx.x]);
    } else {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) dst[(size_t)c * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
  }
}

// One CTA of W/8 threads per object; thread j owns columns 8j..8j+7 (one 16-byte load per row).
template <bool AVG>
__global__ void segpool_bf16_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int col_s, int col_o, int W,
                                    const int* __restrict__ rp_s, const int* __restrict__ perm_s,
                                    const int* __restrict__ rp_o, const int* __restrict__ perm_o,
                                    const int* __restrict__ valid, const float* __restrict__ conf,
                                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int ldo,
                                    float* __restrict__ cnt_out) {
  const int o = blockIdx.x;
  const int c = threadIdx.x * 8;
  const bool colok = c < W;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float cnt = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    const int* rp = pass ? rp_o : rp_s;
    const int* perm = pass ? perm_o : perm_s;
    const int col = (pass ? col_o : col_s) + c;
    const int beg = rp[o], end = rp[o + 1];
    int j = beg;
    for (; j + 4 <= end; j += 4) {
      int t[4];
      bool v[4];
      uint4 r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        t[k] = perm[j + k];
        v[k] = !AVG || valid[t[k]];
        r[k] = (v[k] && colok) ? *reinterpret_cast<const uint4*>(X + (size_t)t[k] * ldx + col) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (v[k]) {
          float f[8];
          unpack8(r[k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
          if (AVG) cnt += conf[t[k]];
        }
      }
    }
    for (; j < end; ++j) {
      int t = perm[j];
      if (AVG && !valid[t]) continue;
      if (colok) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(X + (size_t)t * ldx + col), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
      if (AVG) cnt += conf[t];
    }
  }
  if (AVG && cnt > 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __fdiv_rn(acc[i], cnt);
  }
  if (colok) {
    if (out_f32) {
      st_f4(out_f32 + (size_t)o * ldo + c, make_float4(acc[0], acc[1], acc[2], acc[3]));
      st_f4(out_f32 + (size_t)o * ldo + c + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    if (out_bf16) *reinterpret_cast<uint4*>(out_bf16 + (size_t)o * ldo + c) = pack8(acc);
  }
  if (AVG && threadIdx.x == 0) cnt_out[o] = cnt;
}

__global__ void relu_mask_bf16_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                      __nv_bfloat16* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(__bfloat162float(y[i]) > 0.f ? dy[i] : 0.f);
}

// column sums of a bf16 matrix (bias gradients), fp32 accumulation, ordered two-stage reduction.
// Block = 32 column groups (8 columns = one 16-byte load each) x 8 row lanes, 4 loads in flight per thread.
constexpr int CS_TX = 32, CS_TY = 8, CS_VEC = 8;
__global__ void __launch_bounds__(CS_TX * CS_TY) colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ X, int M, int N,
                                                                           int ld, int rows_per_chunk,
                                                                           float* __restrict__ partial) {
  __shared__ float red[CS_TY][CS_TX * CS_VEC + 4];
  const int tx = threadIdx.x % CS_TX, ty = threadIdx.x / CS_TX;
  const int col = (blockIdx.x * CS_TX + tx) * CS_VEC;
  const int mbeg = blockIdx.y * rows_per_chunk, mend = min(M, mbeg + rows_per_chunk);
  float acc[CS_VEC];
#pragma unroll
  for (int j = 0; j < CS_VEC; ++j) acc[j] = 0.f;
  if (col < N) {
    int m = mbeg + ty;
    for (; m + 3 * CS_TY < mend; m += 4 * CS_TY) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const uint4*>(X + (size_t)(m + u * CS_TY) * ld + col);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        unpack8(q[u], v);
#pragma unroll
        for (int j = 0; j < CS_VEC; ++j) acc[j] += v[j];
      }
    }
    for (; m < mend; m += CS_TY) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(X + (size_t)m * ld + col), v);
#pragma unroll
      for (int j = 0; j < CS_VEC; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < CS_VEC; ++j) red[ty][tx * CS_VEC + j] = acc[j];
  __syncthreads();
  {
    const int n = blockIdx.x * CS_TX * CS_VEC + threadIdx.x;
    if (n < N) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < CS_TY; ++r) s += red[r][threadIdx.x];
      partial[(size_t)blockIdx.y * N + n] = s;
    }
  }
}
__global__ void colsum_bf16_final_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += partial[(size_t)c * N + n];
  out[n] = acc;
}
int colsum_bf16_chunks(int M, int N) {
  int col_blocks = csg_div_up(N, CS_TX * CS_VEC);
  int want = csg_div_up(4 * 148, col_blocks);
  int maxc = csg_div_up(M, 32);
  if (want > maxc) want = maxc;
  return want < 1 ? 1 : want;
}

// bf16 twin of triple_bwd_assemble_kernel (graph.cu): one warp per triple, 8 columns per lane per step.
__global__ void triple_bwd_assemble_bf16_kernel(const __nv_bfloat16* __restrict__ out, const float* __restrict__ dS,
                                                const __nv_bfloat16* __restrict__ d_newp, int ld_newp,
                                                const float* __restrict__ dcnt, const int* __restrict__ s_idx,
                                                const int* __restrict__ o_idx, const int* __restrict__ valid,
                                                const int* __restrict__ type32, const float* __restrict__ conf,
                                                int NT, int H, int Dp, __nv_bfloat16* __restrict__ g,
                                                float* __restrict__ dconf) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= NT) return;
  const int lane = threadIdx.x & 31;
  const int Wd = 2 * H + Dp;
  const int s = s_idx[t], o = o_idx[t];
  const bool v = valid[t] != 0;
  const float cf = conf[t];
  const __nv_bfloat16* orow = out + (size_t)t * Wd;
  __nv_bfloat16* grow = g + (size_t)t * Wd;
  float dot = 0.f;
  for (int j = lane * 8; j < Wd; j += 256) {
    float raw[8];
    if (j < H || j >= H + Dp) {
      const float* src = dS + (size_t)(j < H ? s : o) * H + (j < H ? j : j - H - Dp);
      if (v) {
        float4 a = ld_f4(src), b = ld_f4(src + 4);
        raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = 0.f;
      }
    } else if (d_newp) {
      unpack8(*reinterpret_cast<const uint4*>(d_newp + (size_t)t * ld_newp + (j - H)), raw);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) raw[i] = 0.f;
    }
    float y[8], r[8];
    unpack8(*reinterpret_cast<const uint4*>(orow + j), y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dot += raw[i] * y[i];
      r[i] = y[i] > 0.f ? raw[i] * cf : 0.f;
    }
    *reinterpret_cast<uint4*>(grow + j) = pack8(r);
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    float dc = 0.f;
    if (type32[t] == 1 && cf > 0.f) dc = dot / cf;
    if (v) dc += dcnt[s] + dcnt[o];
    dconf[t] = dc;
  }
}

}  // namespace

CSG_API int csg_cast_bf16(const float* src, int rows, int cols, int ld_src, void* dst, int ld_dst, int transpose,
                          cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  dim3 grid(csg_div_up(cols, 32), csg_div_up(rows, 32));
  cast_bf16_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, rows, cols, ld_src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                     ld_dst, transpose);
  CSG_CHECK_LAUNCH("csg_cast_bf16");
  return 0;
}

CSG_API int csg_segpool_bf16(const void* X, int ldx, int col_s, int col_o, int W,
                             const int* rowptr_s, const int* perm_s, const int* rowptr_o, const int* perm_o,
                             const int* valid, const float* conf, int NO, float* out_f32, void* out_bf16, int ldo,
                             float* cnt_out, int avg, cudaStream_t stream) {
  if (NO == 0) return 0;
  CSG_REQUIRE((W & 7) == 0 && (ldx & 7) == 0 && (col_s & 7) == 0 && (col_o & 7) == 0 && (ldo & 7) == 0,
              "segpool_bf16: widths/offsets must be multiples of 8");
  CSG_REQUIRE(W <= 8192, "segpool_bf16: W=%d too wide", W);
  int threads = ((W / 8 + 31) / 32) * 32;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(X);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (avg) {
    CSG_REQUIRE(valid && conf && cnt_out, "segpool_bf16(avg): valid/conf/cnt required");
    segpool_bf16_kernel<true><<<NO, threads, 0, stream>>>(x, ldx, col_s, col_o, W, rowptr_s, perm_s, rowptr_o, perm_o,
                                                          valid, conf, out_f32, ob, ldo, cnt_out);
  } else {
    segpool_bf16_kernel<false><<<NO, threads, 0, stream>>>(x, ldx, col_s, col_o, W, rowptr_s, perm_s, rowptr_o, perm_o,
                                                           nullptr, nullptr, out_f32, ob, ldo, nullptr);
  }
  CSG_CHECK_LAUNCH("csg_segpool_bf16");
  return 0;
}

CSG_API int csg_relu_mask_bf16(const float* dy, const void* y, void* out, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  relu_mask_bf16_kernel<<<csg_div_up(n, 256), 256, 0, stream>>>(dy, reinterpret_cast<const __nv_bfloat16*>(y),
                                                               reinterpret_cast<__nv_bfloat16*>(out), n);
  CSG_CHECK_LAUNCH("csg_relu_mask_bf16");
  return 0;
}

CSG_API size_t csg_colsum_bf16_workspace(int M, int N) {
  return (size_t)colsum_bf16_chunks(M, N) * N * sizeof(float) + 16;
}

CSG_API int csg_colsum_bf16(const void* X, int M, int N, int ld, float* out, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream) {
  if (N == 0) return 0;
  CSG_REQUIRE((N & 7) == 0 && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
              "colsum_bf16: N and ld must be multiples of 8 and X 16-byte aligned");
  CSG_REQUIRE(workspace_bytes >= csg_colsum_bf16_workspace(M, N), "colsum_bf16: workspace too small");
  const int chunks = colsum_bf16_chunks(M, N);
  const int rows_per_chunk = csg_div_up(M > 0 ? M : 1, chunks);
  float* partial = reinterpret_cast<float*>(workspace);
  colsum_bf16_partial_kernel<<<dim3(csg_div_up(N, CS_TX * CS_VEC), chunks), CS_TX * CS_TY, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(X), M, N, ld, rows_per_chunk, partial);
  CSG_CHECK_LAUNCH("csg_colsum_bf16 partial");
  colsum_bf16_final_kernel<<<csg_div_up(N, 128), 128, 0, stream>>>(partial, chunks, N, out);
  CSG_CHECK_LAUNCH("csg_colsum_bf16 final");
  return 0;
}

CSG_API int csg_triple_bwd_assemble_bf16(const void* out, const float* dS, const void* d_newp, int ld_newp,
                                         const float* dcnt, const int* s_idx, const int* o_idx, const int* valid,
                                         const int* type32, const float* conf, int NT, int H, int Dp, void* g,
                                         float* dconf, cudaStream_t stream) {
  if (NT == 0) return 0;
  CSG_REQUIRE((H & 7) == 0 && (Dp & 7) == 0 && (ld_newp & 7) == 0, "bwd_assemble_bf16: H, Dp, ld must be multiples of 8");
  triple_bwd_assemble_bf16_kernel<<<csg_div_up((long long)NT * 32, 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), dS, reinterpret_cast<const __nv_bfloat16*>(d_newp), ld_newp, dcnt,
      s_idx, o_idx, valid, type32, conf, NT, H, Dp, reinterpret_cast<__nv_bfloat16*>(g), dconf);
  CSG_CHECK_LAUNCH("csg_triple_bwd_assemble_bf16");
  return 0;
}
