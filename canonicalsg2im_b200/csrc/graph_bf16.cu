// bf16-activation variants of the pooling / assembly kernels used by the tensor-core engine, plus
// the fp32 -> bf16 weight cast (with the transposed copy the dX-type GEMMs consume).
// Semantics are those of graph.cu (sg2im/graph.py:69-107); accumulation stays fp32.
#include "common.cuh"
#include "internal.h"
#include <cuda_bf16.h>
#include <stdlib.h>

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
                                 __nv_bfloat16* __restrict__ dst, int ld_dst, int transpose, bool fp16) {
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
      if (r < rows && c < cols) reinterpret_cast<unsigned short*>(dst)[(size_t)r * ld_dst + c] = cvt16(tile[i][threadIdx.x], fp16);
    } else {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) reinterpret_cast<unsigned short*>(dst)[(size_t)c * ld_dst + r] = cvt16(tile[threadIdx.x][i], fp16);
    }
  }
}

// Several fp32 -> bf16 casts (optionally transposed) in one launch: the 4 weight matrices of a layer and their
// transposes are needed once per step each.
constexpr int CAST_MAX = 64;      // 10 copies per GraphTripleConv layer x 6 layers (csg_gconv_bf16_cast_weights)
struct CastJobs {
  const float* src[CAST_MAX];
  __nv_bfloat16* dst[CAST_MAX];
  int rows[CAST_MAX], cols[CAST_MAX], transpose[CAST_MAX];
  int lds[CAST_MAX], ldd[CAST_MAX];   // row pitch of src / dst in elements
  int tile_end[CAST_MAX];      // exclusive prefix of 32x32 tiles
  int n;
};
__global__ void cast_bf16_multi_kernel(const CastJobs jobs, bool fp16) {
  CSG_PDL_WAIT();
  __shared__ float tile[32][33];
  int j = 0;
  while (j < jobs.n - 1 && (int)blockIdx.x >= jobs.tile_end[j]) ++j;
  const int local = blockIdx.x - (j ? jobs.tile_end[j - 1] : 0);
  const int rows = jobs.rows[j], cols = jobs.cols[j];
  const int tcols = (cols + 31) / 32;
  const int c0 = (local % tcols) * 32, r0 = (local / tcols) * 32;
  const float* src = jobs.src[j];
  __nv_bfloat16* dst = jobs.dst[j];
  const int lds = jobs.lds[j], ldd = jobs.ldd[j];
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    if (!jobs.transpose[j]) {
      int r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) reinterpret_cast<unsigned short*>(dst)[(size_t)r * ldd + c] = cvt16(tile[i][threadIdx.x], fp16);
    } else {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) reinterpret_cast<unsigned short*>(dst)[(size_t)c * ldd + r] = cvt16(tile[threadIdx.x][i], fp16);
    }
  }
}

// One CTA per object: TX = W/8 column threads (8 columns = one 16-byte load per row) x TY row lanes.
// The object's incidence list (subject incidences in ascending triple id, then object incidences: the order
// CPU scatter_add visits them, graph.py:98-99) is dealt round-robin to the row lanes, every lane keeps four
// independent row loads in flight, and the lanes are combined in lane order: a fixed, reproducible sum.
constexpr int SP_MAX_THREADS = 256;
constexpr int SP_CHUNK = 256;          // incidences staged per pass
constexpr int SP_UNROLL = 4;           // 16-byte row loads in flight per thread (8 measured slower: 58 vs 53 us)
template <bool AVG>
__global__ void __launch_bounds__(SP_MAX_THREADS)
segpool_bf16_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int col_s, int col_o, int W, int TX, int TY,
                    const int* __restrict__ rp_s, const int* __restrict__ perm_s,
                    const int* __restrict__ rp_o, const int* __restrict__ perm_o,
                    const int* __restrict__ valid, const float* __restrict__ conf,
                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int ldo,
                    float* __restrict__ cnt_out, bool fp16, int split_n) {
  CSG_PDL_WAIT();
  __shared__ float red[SP_MAX_THREADS * 8];
  __shared__ float red_cnt[SP_MAX_THREADS];
  // The incidence list of the object is staged in shared memory SP_CHUNK entries at a time by the whole block (row
  // offset in 16-byte units, confidence; -1 marks a triple that does not take part), so that the row loop issues one
  // shared-memory read per 16-byte row load instead of a perm / valid / conf global load each (ncu r02b: the loop
  // was at 51 % issue-active for 242 MB in 69 us)
  __shared__ unsigned s_off[SP_CHUNK];
  __shared__ float s_w[SP_CHUNK];
  // objects are visited last-to-first: X was just written front-to-back by the producing GEMM and is larger than L2, so
  // its tail is what is still cached; walking forwards would evict that tail before reaching it
  int o = gridDim.x - 1 - blockIdx.x;
  // split_n > 0 (csg_segsum2_bf16): the grid holds 2 * split_n blocks; block 2 o + part sums only the subject (part 0) or
  // only the object (part 1) incidences of object o and writes columns part * W .. of the output row
  // The two blocks of an object are neighbours in launch order: the triples an object takes part in belong to one graph
  // (~1 MB of contiguous rows), so whichever of the 2 x ~18 blocks of a graph comes second finds the rows in L2
  int part = -1;
  if (split_n > 0) { part = o & 1; o >>= 1; }
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int c = tx * 8;
  const bool colok = c < W;
  const int bs = rp_s[o], ns = part == 1 ? 0 : rp_s[o + 1] - bs;
  const int bo = rp_o[o], total = ns + (part == 0 ? 0 : rp_o[o + 1] - bo);
  const int oc = (part == 1 ? W : 0) + c;      // output column of this thread
  const uint4* X16 = reinterpret_cast<const uint4*>(X) + tx;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float cnt = 0.f;
  for (int base = 0; base < total; base += SP_CHUNK) {
    const int n = min(SP_CHUNK, total - base);
    if (base) __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int e = base + i;
      const bool subj = e < ns;
      const int t = subj ? perm_s[bs + e] : perm_o[bo + e - ns];
      s_off[i] = (unsigned)(((size_t)t * ldx + (subj ? col_s : col_o)) >> 3);
      s_w[i] = AVG ? (valid[t] != 0 ? conf[t] : -1.f) : 0.f;
    }
    __syncthreads();
    for (int i0 = ty; i0 < n; i0 += SP_UNROLL * TY) {
      uint4 r[SP_UNROLL];
      float w[SP_UNROLL];
#pragma unroll
      for (int k = 0; k < SP_UNROLL; ++k) {
        const int i = i0 + k * TY;
        w[k] = -1.f;
        r[k] = make_uint4(0, 0, 0, 0);
        if (i < n) {
          w[k] = s_w[i];
          if (colok && !(w[k] < 0.f)) r[k] = X16[s_off[i]];
        }
      }
#pragma unroll
      for (int k = 0; k < SP_UNROLL; ++k) {
        if (!(w[k] < 0.f)) {
          float f[8];
          unpack8_16(r[k], f, fp16);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
          cnt += w[k];
        }
      }
    }
  }
  if (TY > 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(ty * 8 + i) * TX + tx] = acc[i];
    if (tx == 0) red_cnt[ty] = cnt;
    __syncthreads();
    if (ty != 0) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      for (int y = 0; y < TY; ++y) s += red[(y * 8 + i) * TX + tx];
      acc[i] = s;
    }
    cnt = 0.f;
    for (int y = 0; y < TY; ++y) cnt += red_cnt[y];
  }
  if (AVG && cnt > 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __fdiv_rn(acc[i], cnt);
  }
  if (colok) {
    if (out_f32) {
      st_f4(out_f32 + (size_t)o * ldo + oc, make_float4(acc[0], acc[1], acc[2], acc[3]));
      st_f4(out_f32 + (size_t)o * ldo + oc + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    if (out_bf16) *reinterpret_cast<uint4*>(out_bf16 + (size_t)o * ldo + oc) = pack8_16(acc, fp16);
  }
  if (AVG && tx == 0) cnt_out[o] = cnt;
}

__global__ void relu_mask_bf16_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                      __nv_bfloat16* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(pos16(reinterpret_cast<const unsigned short*>(y)[i]) ? dy[i] : 0.f);
}

// column sums of a bf16 matrix (bias gradients), fp32 accumulation, ordered two-stage reduction.
// Block = 32 column groups (8 columns = one 16-byte load each) x 8 row lanes, 4 loads in flight per thread.
constexpr int CS_TX = 32, CS_TY = 8, CS_VEC = 8;
__global__ void __launch_bounds__(CS_TX * CS_TY) colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ X, int M, int N,
                                                                           int ld, int rows_per_chunk,
                                                                           float* __restrict__ partial) {
  CSG_PDL_WAIT();
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
// The same for up to CSM_MAX matrices in one launch (1-D grid, a block finds its job by prefix): the three per-object
// bias gradients of a layer's backward (db4, db3, db1) are ~4 us launches each otherwise.  Same per-block arithmetic and
// the same chunking as the single-matrix kernel, hence bit-identical partials.
constexpr int CSM_MAX = 4;
struct ColsumJobs {
  const __nv_bfloat16* X[CSM_MAX];
  float* partial[CSM_MAX];
  int M[CSM_MAX], N[CSM_MAX], ld[CSM_MAX], rpc[CSM_MAX], xblocks[CSM_MAX], block_end[CSM_MAX];
  int n;
};
__global__ void __launch_bounds__(CS_TX * CS_TY) colsum_bf16_multi_partial_kernel(const ColsumJobs jobs) {
  CSG_PDL_WAIT();
  __shared__ float red[CS_TY][CS_TX * CS_VEC + 4];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.block_end[j]) ++j;
  const int local = (int)blockIdx.x - (j ? jobs.block_end[j - 1] : 0);
  const int bx = local % jobs.xblocks[j], by = local / jobs.xblocks[j];
  const __nv_bfloat16* X = jobs.X[j];
  const int M = jobs.M[j], N = jobs.N[j], ld = jobs.ld[j], rows_per_chunk = jobs.rpc[j];
  const int tx = threadIdx.x % CS_TX, ty = threadIdx.x / CS_TX;
  const int col = (bx * CS_TX + tx) * CS_VEC;
  const int mbeg = by * rows_per_chunk, mend = min(M, mbeg + rows_per_chunk);
  float acc[CS_VEC];
#pragma unroll
  for (int q = 0; q < CS_VEC; ++q) acc[q] = 0.f;
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
        for (int i = 0; i < CS_VEC; ++i) acc[i] += v[i];
      }
    }
    for (; m < mend; m += CS_TY) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(X + (size_t)m * ld + col), v);
#pragma unroll
      for (int i = 0; i < CS_VEC; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < CS_VEC; ++i) red[ty][tx * CS_VEC + i] = acc[i];
  __syncthreads();
  {
    const int n = bx * CS_TX * CS_VEC + threadIdx.x;
    if (n < N) {
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < CS_TY; ++r) sum += red[r][threadIdx.x];
      jobs.partial[j][(size_t)by * N + n] = sum;
    }
  }
}
// final pass: csg_reduce_multi (reduce.cu) with lanes = 8: 32 columns x 8 chunk lanes per block, combined in lane order
int colsum_bf16_chunks(int M, int N) {
  int col_blocks = csg_div_up(N, CS_TX * CS_VEC);
  int want = csg_div_up(4 * 148, col_blocks);
  int maxc = csg_div_up(M, 32);
  if (want > maxc) want = maxc;
  return want < 1 ? 1 : want;
}

// bf16 twin of triple_bwd_assemble_kernel (graph.cu): one warp per triple at a time, 8 columns per lane per step;
// the warps of the (persistent) grid take the triples round-robin.  With CS (column sums requested: the bias
// gradient of net1's second Linear is colsum(g)) every lane also accumulates its 8 x ASM_MAXI columns in
// registers over the warp's triples; the warps of a block are combined in warp order through shared memory and
// one partial row per block is written for the ordered final pass (deterministic for a given grid).
constexpr int ASM_WARPS = 8, ASM_MAXI = 5;      // ASM_MAXI * 256 >= Wd
constexpr int ASM_CTAS_PER_SM = 3;
// HT / DPT: compile-time H and Dp (0 = use the runtime values).  The reference geometry (hidden 512, predicate width
// 128: every layer of the default stack) is instantiated with constants, which folds the segment tests and the
// address arithmetic of the column loop (ncu r02b: the generic loop executes 684 warp instructions per row, a third of
// them index math and branches, and is issue-bound at 49 % issue-active).
template <bool CS, bool OUT_FP16, int HT, int DPT>
__global__ void __launch_bounds__(ASM_WARPS * 32, ASM_CTAS_PER_SM)
triple_bwd_assemble_bf16_kernel(const __nv_bfloat16* __restrict__ out, const float* __restrict__ dS,
                                const __nv_bfloat16* __restrict__ d_newp, int ld_newp,
                                const float* __restrict__ dcnt, const int* __restrict__ s_idx,
                                const int* __restrict__ o_idx, const int* __restrict__ valid,
                                const int* __restrict__ type32, const float* __restrict__ conf,
                                int NT, int H_rt, int Dp_rt, __nv_bfloat16* __restrict__ g,
                                float* __restrict__ dconf, float* __restrict__ cs_partial,
                                const int* __restrict__ pred, int P, float* __restrict__ wt_partial) {
  CSG_PDL_WAIT();
  const int H = HT ? HT : H_rt, Dp = DPT ? DPT : Dp_rt;
  __shared__ __align__(16) float cs_red[CS ? (ASM_WARPS / 2) * ASM_MAXI * 256 : 4];
  // per-warp bins of the confidence gradient by predicate (d w_trans, graph.py:69-74): each warp adds its triples in
  // the order it visits them, the warps are combined in warp order below -> one partial row per block
  extern __shared__ float wt_bins[];       // [ASM_WARPS][P] when wt_partial
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wt_partial) {
    for (int i = threadIdx.x; i < ASM_WARPS * P; i += blockDim.x) wt_bins[i] = 0.f;
    __syncthreads();
  }
  const int Wd = 2 * H + Dp;
  float cs[ASM_MAXI][8];
#pragma unroll
  for (int u = 0; u < ASM_MAXI; ++u)
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[u][i] = 0.f;
  // a block owns a CONTIGUOUS chunk of triples (its warps interleave inside it): consecutive triples belong to the same
  // graph, so the <= ~30 dS rows the chunk gathers (fp32, 2 KB each) stay in this SM's L1 instead of being re-fetched
  // from L2 for every triple (the gathers were as many bytes as the streamed rows)
  const int chunk = (NT + gridDim.x - 1) / gridDim.x;
  const int t_end = min(NT, (int)(blockIdx.x + 1) * chunk);
  const int stride = ASM_WARPS;
  int t = blockIdx.x * chunk + warp;
  // scalars of the next triple are fetched one iteration ahead: the row addresses depend on them
  int s = 0, o = 0, vi = 0, ty = 0, pr = 0;
  float cf = 0.f;
  if (t < t_end) { s = s_idx[t]; o = o_idx[t]; vi = valid[t]; ty = type32[t]; cf = conf[t]; pr = wt_partial ? pred[t] : 0; }
  for (; t < t_end; t += stride) {
    const int tn = t + stride;
    int s2 = 0, o2 = 0, vi2 = 0, ty2 = 0, pr2 = 0;
    float cf2 = 0.f;
    if (tn < t_end) {
      s2 = s_idx[tn]; o2 = o_idx[tn]; vi2 = valid[tn]; ty2 = type32[tn]; cf2 = conf[tn]; pr2 = wt_partial ? pred[tn] : 0;
      // pull the next row of `out` (the only DRAM-resident operand) into L2 while this one is processed (a lead of
      // two rows measured slower: 134 vs 124 us)
      const __nv_bfloat16* nrow = out + (size_t)tn * Wd;
#pragma unroll
      for (int u = 0; u < ASM_MAXI; ++u)
        if (lane * 8 + u * 256 < Wd) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + lane * 8 + u * 256));
    }
    const bool v = vi != 0;
    const __nv_bfloat16* orow = out + (size_t)t * Wd;
    __nv_bfloat16* grow = g + (size_t)t * Wd;
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < ASM_MAXI; ++u) {
      const int j = lane * 8 + u * 256;
      if (j < Wd) {
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
        const uint4 yw = __ldcs(reinterpret_cast<const uint4*>(orow + j));
        unpack8_16(yw, y, OUT_FP16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dot += raw[i] * y[i];
          r[i] = raw[i] * cf;
        }
        // ReLU mask on the packed words: a 16-bit float (bf16 or fp16) is > 0 iff its bits, read as int16, are > 0
        uint4 packed = pack8(r);
        packed.x &= __vcmpgts2(yw.x, 0u); packed.y &= __vcmpgts2(yw.y, 0u);
        packed.z &= __vcmpgts2(yw.z, 0u); packed.w &= __vcmpgts2(yw.w, 0u);
        *reinterpret_cast<uint4*>(grow + j) = packed;
        if (CS) {
          float rb[8];
          unpack8(packed, rb);             // sum what the GEMMs will read: the bf16-rounded values
#pragma unroll
          for (int i = 0; i < 8; ++i) cs[u][i] += rb[i];
        }
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) {
      float dc = 0.f;
      if (ty == 1 && cf > 0.f) dc = dot / cf;
      if (v) dc += dcnt[s] + dcnt[o];
      if (dconf) dconf[t] = dc;
      if (wt_partial && ty == 1) wt_bins[warp * P + pr] += dc;
    }
    s = s2; o = o2; vi = vi2; ty = ty2; cf = cf2; pr = pr2;
  }
  if (wt_partial) {
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < ASM_WARPS; ++w) a += wt_bins[w * P + p];
      wt_partial[(size_t)blockIdx.x * P + p] = a;
    }
  }
  if (CS) {
    // ordered tree over the 8 warps: (w, w+4), then (w, w+2), then (w, w+1); lane owns columns lane*8 + u*256 + i
    for (int half = ASM_WARPS / 2; half >= 1; half >>= 1) {
      if (warp >= half && warp < 2 * half) {
        float* dst = cs_red + (warp - half) * ASM_MAXI * 256;
#pragma unroll
        for (int u = 0; u < ASM_MAXI; ++u) {
          st_f4(dst + u * 256 + lane * 8, make_float4(cs[u][0], cs[u][1], cs[u][2], cs[u][3]));
          st_f4(dst + u * 256 + lane * 8 + 4, make_float4(cs[u][4], cs[u][5], cs[u][6], cs[u][7]));
        }
      }
      __syncthreads();
      if (warp < half) {
        const float* src = cs_red + warp * ASM_MAXI * 256;
#pragma unroll
        for (int u = 0; u < ASM_MAXI; ++u) {
          float4 a = ld_f4(src + u * 256 + lane * 8), b = ld_f4(src + u * 256 + lane * 8 + 4);
          cs[u][0] += a.x; cs[u][1] += a.y; cs[u][2] += a.z; cs[u][3] += a.w;
          cs[u][4] += b.x; cs[u][5] += b.y; cs[u][6] += b.z; cs[u][7] += b.w;
        }
      }
      __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
      for (int u = 0; u < ASM_MAXI; ++u) {
        const int j = lane * 8 + u * 256;
        if (j < Wd) {
          float* dst = cs_partial + (size_t)blockIdx.x * Wd + j;
          st_f4(dst, make_float4(cs[u][0], cs[u][1], cs[u][2], cs[u][3]));
          st_f4(dst + 4, make_float4(cs[u][4], cs[u][5], cs[u][6], cs[u][7]));
        }
      }
    }
  }
}

// ---- pipelined variant (the one the layer executor runs): the rows of `out` -- the only DRAM-resident operand --
// are streamed into a shared-memory ring by one elected producer lane, ASM_WARPS rows (one per consumer warp, contiguous
// in memory because the rows of net1's output are packed) per cp.async.bulk, ASMP_STAGES copies in flight per CTA, and
// the consumer warps read them from shared memory.  The per-row arithmetic, the row -> warp assignment inside a block
// and the reduction orders are those of triple_bwd_assemble_bf16_kernel; what changes is who waits for DRAM: the
// register-staged kernel exposes a DRAM round trip per 16-byte column step of every row (ncu r02b: 36 % warps active,
// 49 % issue-active at 3.7 TB/s).
constexpr int ASMP_STAGES_DEFAULT = 3;      // 3 x 18 KB per CTA: what the ring does not take stays L1 for the gathered dS rows (ncu r02g:
                                    // long-scoreboard stalls on those gathers lead; L1 hit rate 61 % with a 4-stage ring + separate scratch)
constexpr int ASMP_CTAS_PER_SM = 2;
__device__ __forceinline__ uint32_t asmp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void asmp_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void asmp_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void asmp_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void asmp_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) {                  // a protocol bug traps instead of hanging the GPU box
      printf("csg assemble: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void asmp_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

__host__ __device__ inline size_t asmp_ring_bytes(int Wd, int ASMP_STAGES) {
  const size_t ring = (size_t)ASMP_STAGES * ASM_WARPS * Wd * 2, red = (size_t)(ASM_WARPS / 2) * ASM_MAXI * 256 * 4;
  return ((ring > red ? ring : red) + 127) & ~(size_t)127;
}
template <int HT, int DPT, int ASMP_STAGES>
__global__ void __launch_bounds__((ASM_WARPS + 1) * 32, ASMP_CTAS_PER_SM)
triple_bwd_assemble_pipe_kernel(const __nv_bfloat16* __restrict__ out, const float* __restrict__ dS,
                                const __nv_bfloat16* __restrict__ d_newp, int ld_newp,
                                const float* __restrict__ dcnt, const int* __restrict__ s_idx,
                                const int* __restrict__ o_idx, const int* __restrict__ valid,
                                const int* __restrict__ type32, const float* __restrict__ conf,
                                int NT, int H_rt, int Dp_rt, __nv_bfloat16* __restrict__ g,
                                float* __restrict__ cs_partial, const int* __restrict__ pred, int P,
                                float* __restrict__ wt_partial) {
  CSG_PDL_WAIT();
  const int H = HT ? HT : H_rt, Dp = DPT ? DPT : Dp_rt;
  const int Wd = 2 * H + Dp;
  // dynamic shared memory: [ring: ASMP_STAGES x ASM_WARPS rows of Wd bf16][wt_bins][barriers]; the scratch of the final
  // column-sum tree (cs_red, 20 KB) reuses the ring, which is idle by then
  extern __shared__ __align__(128) unsigned char asmp_smem[];
  const uint32_t stage_bytes = (uint32_t)ASM_WARPS * Wd * 2;
  __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(asmp_smem);
  float* cs_red = reinterpret_cast<float*>(asmp_smem);
  float* wt_bins = reinterpret_cast<float*>(asmp_smem + asmp_ring_bytes(Wd, ASMP_STAGES));   // [ASM_WARPS][P]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(wt_bins + ASM_WARPS * P) + 7) & ~(uintptr_t)7);
  const uint32_t full0 = asmp_smem_u32(bars), empty0 = asmp_smem_u32(bars + ASMP_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ASMP_STAGES; ++i) {
      asmp_mbar_init(full0 + 8 * i, 1);
      asmp_mbar_init(empty0 + 8 * i, ASM_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < ASM_WARPS * P; i += blockDim.x) wt_bins[i] = 0.f;
  __syncthreads();

  const int chunk = (NT + gridDim.x - 1) / gridDim.x;
  const int t_beg = min(NT, (int)blockIdx.x * chunk), t_end = min(NT, (int)(blockIdx.x + 1) * chunk);
  const int niter = (t_end - t_beg + ASM_WARPS - 1) / ASM_WARPS;
  float cs[ASM_MAXI][8];
#pragma unroll
  for (int u = 0; u < ASM_MAXI; ++u)
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[u][i] = 0.f;

  if (warp == ASM_WARPS) {
    // ---------------- producer: one bulk copy of up to ASM_WARPS consecutive rows per stage
    if (lane == 0) {
      for (int k = 0; k < niter; ++k) {
        const int st = k % ASMP_STAGES;
        if (k >= ASMP_STAGES) asmp_mbar_wait(empty0 + 8 * st, ((k / ASMP_STAGES) - 1) & 1);
        const int r0 = t_beg + k * ASM_WARPS;
        const uint32_t bytes = (uint32_t)min(ASM_WARPS, t_end - r0) * Wd * 2;
        asmp_mbar_expect_tx(full0 + 8 * st, bytes);
        asmp_bulk_load(asmp_smem_u32(ring) + st * stage_bytes, out + (size_t)r0 * Wd, bytes, full0 + 8 * st);
      }
    }
  } else {
    // ---------------- consumers: warp w takes row w of every stage
    int t = t_beg + warp;
    int s = 0, o = 0, vi = 0, ty = 0, pr = 0;
    float cf = 0.f;
    if (t < t_end) { s = s_idx[t]; o = o_idx[t]; vi = valid[t]; ty = type32[t]; cf = conf[t]; pr = pred[t]; }
    for (int k = 0; k < niter; ++k, t += ASM_WARPS) {
      const int st = k % ASMP_STAGES;
      const int tn = t + ASM_WARPS;
      int s2 = 0, o2 = 0, vi2 = 0, ty2 = 0, pr2 = 0;
      float cf2 = 0.f;
      if (tn < t_end) { s2 = s_idx[tn]; o2 = o_idx[tn]; vi2 = valid[tn]; ty2 = type32[tn]; cf2 = conf[tn]; pr2 = pred[tn]; }
      asmp_mbar_wait(full0 + 8 * st, (k / ASMP_STAGES) & 1);
      if (t < t_end) {
        const bool v = vi != 0;
        const __nv_bfloat16* orow = ring + (size_t)st * (stage_bytes / 2) + (size_t)warp * Wd;
        __nv_bfloat16* grow = g + (size_t)t * Wd;
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < ASM_MAXI; ++u) {
          const int j = lane * 8 + u * 256;
          if (j < Wd) {
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
            const uint4 yw = *reinterpret_cast<const uint4*>(orow + j);
            unpack8(yw, y);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              dot += raw[i] * y[i];
              r[i] = raw[i] * cf;
            }
            uint4 packed = pack8(r);
            packed.x &= __vcmpgts2(yw.x, 0u); packed.y &= __vcmpgts2(yw.y, 0u);
            packed.z &= __vcmpgts2(yw.z, 0u); packed.w &= __vcmpgts2(yw.w, 0u);
            *reinterpret_cast<uint4*>(grow + j) = packed;
            float rb[8];
            unpack8(packed, rb);             // sum what the GEMMs will read: the bf16-rounded values
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[u][i] += rb[i];
          }
        }
        dot = warp_sum(dot);
        if (lane == 0) {
          float dc = 0.f;
          if (ty == 1 && cf > 0.f) dc = dot / cf;
          if (v) dc += dcnt[s] + dcnt[o];
          if (ty == 1) wt_bins[warp * P + pr] += dc;
        }
      }
      __syncwarp();
      if (lane == 0) asmp_mbar_arrive(empty0 + 8 * st);      // the row has been read: the stage may be refilled
      s = s2; o = o2; vi = vi2; ty = ty2; cf = cf2; pr = pr2;
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < ASM_WARPS; ++w) a += wt_bins[w * P + p];
    wt_partial[(size_t)blockIdx.x * P + p] = a;
  }
  // ordered tree over the 8 consumer warps: (w, w+4), then (w, w+2), then (w, w+1); lane owns columns lane*8 + u*256 + i
  for (int half = ASM_WARPS / 2; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
      float* dst = cs_red + (warp - half) * ASM_MAXI * 256;
#pragma unroll
      for (int u = 0; u < ASM_MAXI; ++u) {
        st_f4(dst + u * 256 + lane * 8, make_float4(cs[u][0], cs[u][1], cs[u][2], cs[u][3]));
        st_f4(dst + u * 256 + lane * 8 + 4, make_float4(cs[u][4], cs[u][5], cs[u][6], cs[u][7]));
      }
    }
    __syncthreads();
    if (warp < half) {
      const float* src = cs_red + warp * ASM_MAXI * 256;
#pragma unroll
      for (int u = 0; u < ASM_MAXI; ++u) {
        float4 a = ld_f4(src + u * 256 + lane * 8), b = ld_f4(src + u * 256 + lane * 8 + 4);
        cs[u][0] += a.x; cs[u][1] += a.y; cs[u][2] += a.z; cs[u][3] += a.w;
        cs[u][4] += b.x; cs[u][5] += b.y; cs[u][6] += b.z; cs[u][7] += b.w;
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
#pragma unroll
    for (int u = 0; u < ASM_MAXI; ++u) {
      const int j = lane * 8 + u * 256;
      if (j < Wd) {
        float* dst = cs_partial + (size_t)blockIdx.x * Wd + j;
        st_f4(dst, make_float4(cs[u][0], cs[u][1], cs[u][2], cs[u][3]));
        st_f4(dst + 4, make_float4(cs[u][4], cs[u][5], cs[u][6], cs[u][7]));
      }
    }
  }
}
size_t asmp_smem_bytes(int Wd, int P, int stages) {
  return asmp_ring_bytes(Wd, stages) + (size_t)ASM_WARPS * P * 4 + 8 + 2 * stages * 8;
}
int asmp_blocks(int NT) {
  int want = csg_div_up(NT > 0 ? NT : 1, ASM_WARPS);
  int cap = csg_num_sms() * ASMP_CTAS_PER_SM;
  return want < cap ? want : cap;
}

int asm_blocks(int NT) {
  int want = csg_div_up(NT > 0 ? NT : 1, ASM_WARPS);
  int cap = csg_num_sms() * ASM_CTAS_PER_SM;
  return want < cap ? want : cap;
}

}  // namespace

CSG_API int csg_cast_bf16(const float* src, int rows, int cols, int ld_src, void* dst, int ld_dst, int transpose, int fp16,
                          cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  dim3 grid(csg_div_up(cols, 32), csg_div_up(rows, 32));
  cast_bf16_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, rows, cols, ld_src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                     ld_dst, transpose, fp16 != 0);
  CSG_CHECK_LAUNCH("csg_cast_bf16");
  return 0;
}

// n contiguous fp32 matrices src[i] [rows[i], cols[i]] -> contiguous bf16 dst[i] ([cols, rows] when transpose[i]).
// The pointer / size arrays are HOST arrays of length n <= 64.
CSG_API int csg_cast_bf16_multi(int n, const void* const* src, void* const* dst, const int* rows, const int* cols,
                                const int* transpose, int fp16, cudaStream_t stream) {
  return csg_cast_bf16_multi_ld(n, src, dst, rows, cols, transpose, nullptr, nullptr, fp16, stream);
}

// the same with row pitches (elements; NULL or 0 = contiguous): sub-matrices of the weights, column slices of a wider copy
int csg_cast_bf16_multi_ld(int n, const void* const* src, void* const* dst, const int* rows, const int* cols,
                           const int* transpose, const int* ld_src, const int* ld_dst, int fp16, cudaStream_t stream) {
  if (n == 0) return 0;
  CSG_REQUIRE(n > 0 && n <= CAST_MAX, "cast_bf16_multi: n=%d out of range", n);
  CastJobs jobs;
  jobs.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    jobs.src[i] = reinterpret_cast<const float*>(src[i]);
    jobs.dst[i] = reinterpret_cast<__nv_bfloat16*>(dst[i]);
    jobs.rows[i] = rows[i]; jobs.cols[i] = cols[i]; jobs.transpose[i] = transpose[i];
    jobs.lds[i] = (ld_src && ld_src[i]) ? ld_src[i] : cols[i];
    jobs.ldd[i] = (ld_dst && ld_dst[i]) ? ld_dst[i] : (transpose[i] ? rows[i] : cols[i]);
    total += csg_div_up(rows[i], 32) * csg_div_up(cols[i], 32);
    jobs.tile_end[i] = total;
  }
  if (total == 0) return 0;
  CSG_CUDA(csg_launch_pdl(cast_bf16_multi_kernel, dim3(total), dim3(dim3(32, 8)), 0, stream, jobs, fp16 != 0));
  CSG_CHECK_LAUNCH("csg_cast_bf16_multi");
  return 0;
}

CSG_API int csg_segpool_bf16(const void* X, int ldx, int col_s, int col_o, int W,
                             const int* rowptr_s, const int* perm_s, const int* rowptr_o, const int* perm_o,
                             const int* valid, const float* conf, int NO, float* out_f32, void* out_bf16, int ldo,
                             float* cnt_out, int avg, int fp16, cudaStream_t stream) {
  if (NO == 0) return 0;
  CSG_REQUIRE((W & 7) == 0 && (ldx & 7) == 0 && (col_s & 7) == 0 && (col_o & 7) == 0 && (ldo & 7) == 0,
              "segpool_bf16: widths/offsets must be multiples of 8");
  CSG_REQUIRE(W <= 8 * SP_MAX_THREADS, "segpool_bf16: W=%d too wide", W);
  const int TX = W / 8;
  int TY = SP_MAX_THREADS / TX;
  if (TY > 16) TY = 16;
  if (TY < 1) TY = 1;
  const int threads = TX * TY;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(X);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (avg) {
    // valid / conf are read per incidence only: a batch without triples may pass NULL for them
    CSG_REQUIRE(cnt_out, "segpool_bf16(avg): cnt_out required");
    CSG_CUDA(csg_launch_pdl(segpool_bf16_kernel<true>, dim3(NO), dim3(threads), 0, stream, x, ldx, col_s, col_o, W, TX, TY, rowptr_s, perm_s, rowptr_o,
                                                          perm_o, valid, conf, out_f32, ob, ldo, cnt_out, fp16 != 0, 0));
  } else {
    CSG_CUDA(csg_launch_pdl(segpool_bf16_kernel<false>, dim3(NO), dim3(threads), 0, stream, x, ldx, col_s, col_o, W, TX, TY, rowptr_s, perm_s, rowptr_o,
                                                           perm_o, nullptr, nullptr, out_f32, ob, ldo, nullptr, fp16 != 0, 0));
  }
  CSG_CHECK_LAUNCH("csg_segpool_bf16");
  return 0;
}

// out[o, 0:W] = sum over the triples whose SUBJECT is o of X[t, 0:W]; out[o, W:2W] = the same over the triples whose
// OBJECT is o (ascending triple id inside a row: fixed summation order).  One launch, 2 * NO blocks.
CSG_API int csg_segsum2_bf16(const void* X, int ldx, int W, const int* rowptr_s, const int* perm_s, const int* rowptr_o,
                             const int* perm_o, int NO, float* out_f32, void* out_bf16, int ldo, cudaStream_t stream) {
  if (NO == 0) return 0;
  CSG_REQUIRE((W & 7) == 0 && (ldx & 7) == 0 && (ldo & 7) == 0 && ldo >= 2 * W, "segsum2_bf16: W, ldx, ldo must be multiples of 8 and ldo >= 2 W");
  CSG_REQUIRE(W <= 8 * SP_MAX_THREADS, "segsum2_bf16: W=%d too wide", W);
  CSG_REQUIRE(out_f32 || out_bf16, "segsum2_bf16: no output");
  const int TX = W / 8;
  int TY = SP_MAX_THREADS / TX;
  if (TY > 16) TY = 16;
  if (TY < 1) TY = 1;
  CSG_CUDA(csg_launch_pdl(segpool_bf16_kernel<false>, dim3(2 * NO), dim3(TX * TY), 0, stream, reinterpret_cast<const __nv_bfloat16*>(X), ldx, 0, 0, W,
                          TX, TY, rowptr_s, perm_s, rowptr_o, perm_o, nullptr, nullptr, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                          ldo, nullptr, false, NO));
  CSG_CHECK_LAUNCH("csg_segsum2_bf16");
  return 0;
}

// ---- fp32 accuracy on the tensor cores: x = hi + mid + lo with three bf16 terms (8 + 8 + 8 = the 24 significant bits of
// fp32, so the split is exact outside the subnormal range) and
//     a b  ~=  a_hi b_hi + a_hi b_mid + a_mid b_hi + a_hi b_lo + a_lo b_hi + a_mid b_mid
// (the three dropped products are below 2^-24 |a b|; every kept product of two bf16 numbers is exact in fp32 and the sums
// run in the fp32 accumulators of tcgen05.mma).  Laid out along the reduction dimension this is ONE bf16 GEMM with a 6x
// longer K: the A operand carries its parts in the order (mid, lo, hi, mid, hi, hi), the B operand (mid, hi, lo, hi, mid, hi).
// csg_split3_bf16 writes that K-concatenated operand: reduction dimension = columns (K-major operand, out [rows, 6 cols])
// or = rows (MN-major operand, out [6 rows, cols]); `transpose` first transposes the fp32 input (weights given as [K, N]).
namespace {
__global__ void split3_bf16_kernel(const float* __restrict__ X, int rows, int cols, int ld, int transpose, int k_is_cols,
                                   int role, __nv_bfloat16* __restrict__ out, int ld_out) {
  CSG_PDL_WAIT();
  __shared__ float tile[32][33];
  // logical matrix L [R, C] = X (or X^T); tiles of 32 x 32 through shared memory so that reads and writes stay coalesced
  const int R = transpose ? cols : rows, C = transpose ? rows : cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    float v = 0.f;
    if (!transpose) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      if (r < R && c < C) v = X[(size_t)r * ld + c];
      tile[i][threadIdx.x] = v;
    } else {
      // L[r, c] = X[c, r]: read X rows c0+i (= L columns), contiguous along r
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) v = X[(size_t)c * ld + r];
      tile[threadIdx.x][i] = v;
    }
  }
  __syncthreads();
  // order of the six products along K: smallest first (mid*mid, lo*hi, hi*lo ~ 2^-16; mid*hi, hi*mid ~ 2^-8), hi*hi LAST.
  // tcgen05.mma truncates its fp32 accumulator at every instruction, by up to an ulp of the RUNNING sum: while the
  // corrections accumulate the running sum is 2^-8 .. 2^-16 of the result, so only the K/16 instructions of the hi*hi
  // block truncate at full scale (measured: 6x less error than with hi*hi first)
  const int pa[6] = {1, 2, 0, 1, 0, 0}, pb[6] = {1, 0, 2, 0, 1, 0};
  // role: 0 / 1 = all six blocks of the A / B operand; 2 / 3 = the hi block alone; 4 / 5 = the five correction blocks
  // (weight gradients reduce over ~1e5 rows: their hi*hi products are run as short chains of their own, ops.py)
  const bool is_b = role & 1;
  const int kbeg = (role >> 1) == 1 ? 5 : 0, kend = (role >> 1) == 2 ? 5 : 6;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r >= R || c >= C) continue;
    const float x = tile[i][threadIdx.x];
    __nv_bfloat16 part[3];
    part[0] = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(part[0]);
    part[1] = __float2bfloat16_rn(r1);
    part[2] = __float2bfloat16_rn(r1 - __bfloat162float(part[1]));
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k < kbeg || k >= kend) continue;
      const __nv_bfloat16 v = part[is_b ? pb[k] : pa[k]];
      const int kk = k - kbeg;
      if (k_is_cols) out[(size_t)r * ld_out + (size_t)kk * C + c] = v;
      else out[((size_t)kk * R + r) * ld_out + c] = v;
    }
  }
}
}  // namespace

CSG_API int csg_split3_bf16(const float* X, int rows, int cols, int ld, int transpose, int k_is_cols, int role, void* out,
                            int ld_out, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  CSG_REQUIRE(role >= 0 && role <= 5, "split3_bf16: role must be 0..5");
  const int R = transpose ? cols : rows, C = transpose ? rows : cols;
  const int nblk = (role >> 1) == 0 ? 6 : ((role >> 1) == 1 ? 1 : 5);
  CSG_REQUIRE(ld_out >= (k_is_cols ? nblk * C : C), "split3_bf16: ld_out too small");
  dim3 grid(csg_div_up(C, 32), csg_div_up(R, 32));
  CSG_CUDA(csg_launch_pdl(split3_bf16_kernel, grid, dim3(32, 8), 0, stream, X, rows, cols, ld, transpose, k_is_cols, role,
                          reinterpret_cast<__nv_bfloat16*>(out), ld_out));
  CSG_CHECK_LAUNCH("csg_split3_bf16");
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
  CsgReduceJob job;
  if (int rc = csg_colsum_bf16_deferred(X, M, N, ld, out, workspace, workspace_bytes, stream, &job)) return rc;
  return csg_reduce_multi(&job, 1, stream);
}

int csg_colsum_bf16_deferred(const void* X, int M, int N, int ld, float* out, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream, CsgReduceJob* job) {
  job->parts = 0; job->n = 0; job->partial = nullptr; job->out = out; job->stride = N; job->lanes = 8;
  job->op = CSG_RED_SUM; job->aux = nullptr; job->ncols = 0; job->ldo = 0;
  if (N == 0) return 0;
  CSG_REQUIRE((N & 7) == 0 && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
              "colsum_bf16: N and ld must be multiples of 8 and X 16-byte aligned");
  CSG_REQUIRE(workspace_bytes >= csg_colsum_bf16_workspace(M, N), "colsum_bf16: workspace too small");
  const int chunks = colsum_bf16_chunks(M, N);
  const int rows_per_chunk = csg_div_up(M > 0 ? M : 1, chunks);
  float* partial = reinterpret_cast<float*>(workspace);
  CSG_CUDA(csg_launch_pdl(colsum_bf16_partial_kernel, dim3(dim3(csg_div_up(N, CS_TX * CS_VEC), chunks)), dim3(CS_TX * CS_TY), 0, stream,
      reinterpret_cast<const __nv_bfloat16*>(X), M, N, ld, rows_per_chunk, partial));
  CSG_CHECK_LAUNCH("csg_colsum_bf16 partial");
  job->partial = partial; job->n = N; job->parts = chunks;
  return 0;
}

// n <= 4 column sums in one launch; job[i] describes the final pass of matrix i (as csg_colsum_bf16_deferred)
int csg_colsum_bf16_multi_deferred(int n, const void* const* X, const int* M, const int* N, const int* ld, float* const* out,
                                   void* const* workspace, const size_t* workspace_bytes, cudaStream_t stream,
                                   CsgReduceJob* job) {
  CSG_REQUIRE(n >= 0 && n <= CSM_MAX, "colsum_bf16_multi: n=%d out of range", n);
  ColsumJobs jobs;
  jobs.n = 0;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    job[i].parts = 0; job[i].n = 0; job[i].partial = nullptr; job[i].out = out[i]; job[i].stride = N[i]; job[i].lanes = 8;
    job[i].op = CSG_RED_SUM; job[i].aux = nullptr; job[i].ncols = 0; job[i].ldo = 0;
    if (N[i] == 0) continue;
    CSG_REQUIRE((N[i] & 7) == 0 && (ld[i] & 7) == 0 && (reinterpret_cast<uintptr_t>(X[i]) & 15) == 0,
                "colsum_bf16: N and ld must be multiples of 8 and X 16-byte aligned");
    CSG_REQUIRE(workspace_bytes[i] >= csg_colsum_bf16_workspace(M[i], N[i]), "colsum_bf16: workspace too small");
    const int chunks = colsum_bf16_chunks(M[i], N[i]);
    const int k = jobs.n++;
    jobs.X[k] = reinterpret_cast<const __nv_bfloat16*>(X[i]);
    jobs.partial[k] = reinterpret_cast<float*>(workspace[i]);
    jobs.M[k] = M[i]; jobs.N[k] = N[i]; jobs.ld[k] = ld[i];
    jobs.rpc[k] = csg_div_up(M[i] > 0 ? M[i] : 1, chunks);
    jobs.xblocks[k] = csg_div_up(N[i], CS_TX * CS_VEC);
    blocks += jobs.xblocks[k] * chunks;
    jobs.block_end[k] = blocks;
    job[i].partial = jobs.partial[k]; job[i].n = N[i]; job[i].parts = chunks;
  }
  if (jobs.n == 0) return 0;
  for (int k = jobs.n; k < CSM_MAX; ++k) {
    jobs.X[k] = jobs.X[0]; jobs.partial[k] = jobs.partial[0]; jobs.M[k] = 0; jobs.N[k] = 0; jobs.ld[k] = 0; jobs.rpc[k] = 1;
    jobs.xblocks[k] = 1; jobs.block_end[k] = blocks;
  }
  CSG_CUDA(csg_launch_pdl(colsum_bf16_multi_partial_kernel, dim3(blocks), dim3(CS_TX * CS_TY), 0, stream, jobs));
  CSG_CHECK_LAUNCH("csg_colsum_bf16 multi partial");
  return 0;
}

CSG_API size_t csg_triple_bwd_assemble_bf16_workspace(int NT, int H, int Dp) {
  const int Wd = 2 * H + Dp;
  size_t fused = (size_t)asm_blocks(NT) * Wd * sizeof(float) + 16;
  size_t plain = csg_colsum_bf16_workspace(NT, Wd);
  return fused > plain ? fused : plain;
}

// colsum_g (may be null): [2H+Dp] fp32 column sums of g, i.e. the bias gradient of net1's second Linear.
CSG_API int csg_triple_bwd_assemble_bf16(const void* out, const float* dS, const void* d_newp, int ld_newp,
                                         const float* dcnt, const int* s_idx, const int* o_idx, const int* valid,
                                         const int* type32, const float* conf, int NT, int H, int Dp, void* g,
                                         float* dconf, float* colsum_g, int out_fp16, void* workspace,
                                         size_t workspace_bytes, cudaStream_t stream) {
  const int Wd = 2 * H + Dp;
  if (NT == 0) {
    if (colsum_g) CSG_CUDA(cudaMemsetAsync(colsum_g, 0, (size_t)Wd * sizeof(float), stream));
    return 0;
  }
  CSG_REQUIRE((H & 7) == 0 && (Dp & 7) == 0 && (ld_newp & 7) == 0, "bwd_assemble_bf16: H, Dp, ld must be multiples of 8");
  CSG_REQUIRE(Wd <= ASM_MAXI * 256, "bwd_assemble_bf16: 2H+Dp=%d exceeds %d", Wd, ASM_MAXI * 256);
  const int blocks = asm_blocks(NT);
  const __nv_bfloat16* o16 = reinterpret_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* p16 = reinterpret_cast<const __nv_bfloat16*>(d_newp);
  __nv_bfloat16* g16 = reinterpret_cast<__nv_bfloat16*>(g);
  if (colsum_g) {
    CSG_REQUIRE(workspace && workspace_bytes >= csg_triple_bwd_assemble_bf16_workspace(NT, H, Dp),
                "bwd_assemble_bf16: workspace too small");
    float* partial = reinterpret_cast<float*>(workspace);
    if (out_fp16) {
      CSG_CUDA(csg_launch_pdl(triple_bwd_assemble_bf16_kernel<true, true, 0, 0>, dim3(blocks), dim3(ASM_WARPS * 32), 0, stream,
          o16, dS, p16, ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, g16, dconf, partial,
          (const int*)nullptr, 0, (float*)nullptr));
    } else {
      CSG_CUDA(csg_launch_pdl(triple_bwd_assemble_bf16_kernel<true, false, 0, 0>, dim3(blocks), dim3(ASM_WARPS * 32), 0, stream,
          o16, dS, p16, ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, g16, dconf, partial,
          (const int*)nullptr, 0, (float*)nullptr));
    }
    CSG_CHECK_LAUNCH("csg_triple_bwd_assemble_bf16");
    CsgReduceJob job = {partial, colsum_g, Wd, blocks, (long long)Wd, 8, CSG_RED_SUM, nullptr};
    if (int rc = csg_reduce_multi(&job, 1, stream)) return rc;
  } else {
    if (out_fp16) {
      CSG_CUDA(csg_launch_pdl(triple_bwd_assemble_bf16_kernel<false, true, 0, 0>, dim3(blocks), dim3(ASM_WARPS * 32), 0, stream,
          o16, dS, p16, ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, g16, dconf, (float*)nullptr,
          (const int*)nullptr, 0, (float*)nullptr));
    } else {
      CSG_CUDA(csg_launch_pdl(triple_bwd_assemble_bf16_kernel<false, false, 0, 0>, dim3(blocks), dim3(ASM_WARPS * 32), 0, stream,
          o16, dS, p16, ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, g16, dconf, (float*)nullptr,
          (const int*)nullptr, 0, (float*)nullptr));
    }
    CSG_CHECK_LAUNCH("csg_triple_bwd_assemble_bf16");
  }
  return 0;
}

size_t csg_triple_bwd_assemble_bf16_deferred_workspace(int NT, int H, int Dp, int P) {
  return (size_t)asm_blocks(NT) * (2 * H + Dp + P) * sizeof(float) + 64;
}

int csg_triple_bwd_assemble_bf16_deferred(const void* out, const float* dS, const void* d_newp, int ld_newp,
                                          const float* dcnt, const int* s_idx, const int* o_idx, const int* valid,
                                          const int* type32, const int* pred, const float* conf, const float* w_trans,
                                          int NT, int H, int Dp, int P, void* g, float* db2, float* dwt, int out_fp16,
                                          void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                          CsgReduceJob* job_db2, CsgReduceJob* job_dwt) {
  const int Wd = 2 * H + Dp;
  *job_db2 = CsgReduceJob{nullptr, db2, 0, 0, (long long)Wd, 8, CSG_RED_SUM, nullptr};
  *job_dwt = CsgReduceJob{nullptr, dwt, 0, 0, (long long)P, 8, CSG_RED_SIGMOID_GRAD, w_trans};
  if (NT == 0) {
    CSG_CUDA(cudaMemsetAsync(db2, 0, (size_t)Wd * sizeof(float), stream));
    CSG_CUDA(cudaMemsetAsync(dwt, 0, (size_t)P * sizeof(float), stream));
    return 0;
  }
  CSG_REQUIRE((H & 7) == 0 && (Dp & 7) == 0 && (ld_newp & 7) == 0, "bwd_assemble_bf16: H, Dp, ld must be multiples of 8");
  CSG_REQUIRE(Wd <= ASM_MAXI * 256, "bwd_assemble_bf16: 2H+Dp=%d exceeds %d", Wd, ASM_MAXI * 256);
  CSG_REQUIRE(P > 0 && (size_t)ASM_WARPS * P * sizeof(float) <= 24 * 1024, "bwd_assemble_bf16: P=%d predicates do not fit the per-warp bins", P);
  CSG_REQUIRE(workspace && workspace_bytes >= csg_triple_bwd_assemble_bf16_deferred_workspace(NT, H, Dp, P),
              "bwd_assemble_bf16: workspace too small");
  CSG_REQUIRE(!out_fp16, "bwd_assemble_bf16 (deferred): fp16 forward tensors are inference-only");
  {
    // pipelined kernel (rows of `out` through a shared-memory ring); CSG_ASM_PIPE=0 selects the register-staged one
    const char* e = getenv("CSG_ASM_PIPE");
    const char* es = getenv("CSG_ASM_STAGES");
    const int stages = (es && (atoi(es) == 4 || atoi(es) == 2)) ? atoi(es) : ASMP_STAGES_DEFAULT;
    const size_t smem = asmp_smem_bytes(Wd, P, stages);
    if (!(e && e[0] == '0') && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && smem <= 100 * 1024) {
      const int blocks = asmp_blocks(NT);
      float* cs_partial = reinterpret_cast<float*>(workspace);
      float* wt_partial = cs_partial + (size_t)blocks * Wd;
      const bool ref_geom = H == 512 && Dp == 128;
      auto kernel = stages == 4 ? (ref_geom ? triple_bwd_assemble_pipe_kernel<512, 128, 4> : triple_bwd_assemble_pipe_kernel<0, 0, 4>)
                  : stages == 2 ? (ref_geom ? triple_bwd_assemble_pipe_kernel<512, 128, 2> : triple_bwd_assemble_pipe_kernel<0, 0, 2>)
                                : (ref_geom ? triple_bwd_assemble_pipe_kernel<512, 128, 3> : triple_bwd_assemble_pipe_kernel<0, 0, 3>);
      static bool configured[6] = {false, false, false, false, false, false};
      const int which = (ref_geom ? 0 : 1) + (stages == 4 ? 2 : (stages == 2 ? 4 : 0));
      if (!configured[which]) {
        CSG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured[which] = true;
      }
      CSG_CUDA(csg_launch_pdl(kernel, dim3(blocks), dim3((ASM_WARPS + 1) * 32), smem, stream,
                              reinterpret_cast<const __nv_bfloat16*>(out), dS, reinterpret_cast<const __nv_bfloat16*>(d_newp),
                              ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, reinterpret_cast<__nv_bfloat16*>(g),
                              cs_partial, pred, P, wt_partial));
      CSG_CHECK_LAUNCH("csg_triple_bwd_assemble_bf16 (pipelined)");
      job_db2->partial = cs_partial; job_db2->n = Wd; job_db2->parts = blocks;
      job_dwt->partial = wt_partial; job_dwt->n = P; job_dwt->parts = blocks;
      return 0;
    }
  }
  const int blocks = asm_blocks(NT);
  float* cs_partial = reinterpret_cast<float*>(workspace);
  float* wt_partial = cs_partial + (size_t)blocks * Wd;
  auto kernel = (H == 512 && Dp == 128) ? triple_bwd_assemble_bf16_kernel<true, false, 512, 128>
                                        : triple_bwd_assemble_bf16_kernel<true, false, 0, 0>;
  CSG_CUDA(csg_launch_pdl(kernel, dim3(blocks), dim3(ASM_WARPS * 32), (size_t)ASM_WARPS * P * sizeof(float),
                          stream, reinterpret_cast<const __nv_bfloat16*>(out), dS, reinterpret_cast<const __nv_bfloat16*>(d_newp),
                          ld_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, reinterpret_cast<__nv_bfloat16*>(g),
                          (float*)nullptr, cs_partial, pred, P, wt_partial));
  CSG_CHECK_LAUNCH("csg_triple_bwd_assemble_bf16 (deferred)");
  job_db2->partial = cs_partial; job_db2->n = Wd; job_db2->parts = blocks;
  job_dwt->partial = wt_partial; job_dwt->n = P; job_dwt->parts = blocks;
  return 0;
}
