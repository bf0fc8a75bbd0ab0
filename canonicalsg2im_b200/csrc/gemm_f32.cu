// fp32 SIMT GEMM family (the "parity" precision of the triple / object MLPs).
//
// C[M, N] = epilogue(op(A) * op(B)) with fp32 FMA accumulation, used for net1 / net2 /
// box_net (sg2im/graph.py:33-41, sg2im/model.py:58-60) forward and backward when results
// have to agree with the reference's fp32 path to 1e-5.  The tensor-core path
// (gemm_tc.cu) shares the same operand conventions.
//
// Operand modes
//   A_ROW    A[m * lda + k]                 activations [M, K]
//   A_COL    A[k * lda + m]                 activations transposed (weight gradients, K = rows)
//   A_GATHER row m = [obj[s_idx[m]] | pred[m] | obj[o_idx[m]]]   (graph.py:63-66 fused: the
//            concatenated triple input is never materialised)
//   B_NK     B[n * ldb + k]                 nn.Linear weight [out, in]      (y = x W^T)
//   B_KN     B[k * ldb + n]                 weight used as [K, N]           (dx = dy W), or
//                                           activations for weight gradients (dW = dy^T x)
//   B_GATHER row k = gathered triple input  (dW1 = dh^T x with x gathered on the fly)
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

enum { A_ROW = 0, A_COL = 1, A_GATHER = 2 };
enum { B_NK = 0, B_KN = 1, B_GATHER = 2 };

struct Gather {
  const float* obj;    // [NO, Din]
  const float* pred;   // [NT, Dp]
  const int* s_idx;    // [NT]
  const int* o_idx;    // [NT]
  int Din, Dp;
  int ldp;             // row stride of pred (it may be a column slice of the previous layer's net1 output)
};

struct GemmParams {
  const float* A;
  const float* B;
  float* C;
  int M, N, K, lda, ldb, ldc;
  const float* bias;       // [N] or null
  int relu;
  const float* rowscale;   // [M] or null  (triple confidence, graph.py:76-77)
  const float* mask_aux;   // [M, ld_aux] or null: multiply by (aux > 0)  (ReLU backward)
  int ld_aux;
  Gather g;
  int k_split;             // K range per blockIdx.z (split-K writes raw partials [z][M][N])
};

__device__ __forceinline__ float4 gather4(const Gather& g, int t, int c) {
  const float* src;
  if (c < g.Din) src = g.obj + (size_t)g.s_idx[t] * g.Din + c;
  else if (c < g.Din + g.Dp) src = g.pred + (size_t)t * g.ldp + (c - g.Din);
  else src = g.obj + (size_t)g.o_idx[t] * g.Din + (c - g.Din - g.Dp);
  return ld_f4(src);
}

// Loads the 8 elements a thread contributes to a [BK x 128] operand tile.
// "row-like" modes (ROW / NK / GATHER-A): thread -> (mn = tid >> 1, k8 = (tid & 1) * 8), contiguous along k.
// "col-like" modes (COL / KN / GATHER-B): thread -> (k = tid >> 4, mn8 = (tid & 15) * 8), contiguous along mn.
template <int MODE, bool IS_A>
__device__ __forceinline__ void load_tile(const GemmParams& p, int mn0, int k0, int kend, float4& v0, float4& v1) {
  const int tid = threadIdx.x;
  const float* base = IS_A ? p.A : p.B;
  const int ld = IS_A ? p.lda : p.ldb;
  const int MN = IS_A ? p.M : p.N;
  v0 = make_float4(0.f, 0.f, 0.f, 0.f);
  v1 = v0;
  constexpr bool rowlike = IS_A ? (MODE == A_ROW || MODE == A_GATHER) : (MODE == B_NK);
  if (rowlike) {
    int mn = mn0 + (tid >> 1), k = k0 + (tid & 1) * 8;
    if (mn < MN) {
      if (IS_A && MODE == A_GATHER) {
        if (k < kend) v0 = gather4(p.g, mn, k);
        if (k + 4 < kend) v1 = gather4(p.g, mn, k + 4);
      } else {
        const float* src = base + (size_t)mn * ld + k;
        if (k < kend) v0 = ld_f4(src);
        if (k + 4 < kend) v1 = ld_f4(src + 4);
      }
    }
  } else {
    int k = k0 + (tid >> 4), mn = mn0 + (tid & 15) * 8;
    if (k < kend) {
      if (!IS_A && MODE == B_GATHER) {
        if (mn < MN) v0 = gather4(p.g, k, mn);
        if (mn + 4 < MN) v1 = gather4(p.g, k, mn + 4);
      } else {
        const float* src = base + (size_t)k * ld + mn;
        if (mn + 3 < MN) v0 = ld_f4(src);
        else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4 && mn + i < MN; ++i) t[i] = src[i]; v0 = make_float4(t[0], t[1], t[2], t[3]); }
        if (mn + 7 < MN) v1 = ld_f4(src + 4);
        else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4 && mn + 4 + i < MN; ++i) t[i] = src[4 + i]; v1 = make_float4(t[0], t[1], t[2], t[3]); }
      }
    }
  }
}

template <int MODE, bool IS_A>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float4& v0, const float4& v1) {
  const int tid = threadIdx.x;
  constexpr bool rowlike = IS_A ? (MODE == A_ROW || MODE == A_GATHER) : (MODE == B_NK);
  if (rowlike) {
    int mn = tid >> 1, k = (tid & 1) * 8;
    S[k + 0][mn] = v0.x; S[k + 1][mn] = v0.y; S[k + 2][mn] = v0.z; S[k + 3][mn] = v0.w;
    S[k + 4][mn] = v1.x; S[k + 5][mn] = v1.y; S[k + 6][mn] = v1.z; S[k + 7][mn] = v1.w;
  } else {
    int k = tid >> 4, mn = (tid & 15) * 8;
    st_f4(&S[k][mn], v0);
    st_f4(&S[k][mn + 4], v1);
  }
}

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_split;
  const int kend = min(p.K, kbeg + p.k_split);
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 a0, a1, b0, b1;
  load_tile<AMODE, true>(p, m0, kbeg, kend, a0, a1);
  load_tile<BMODE, false>(p, n0, kbeg, kend, b0, b1);
  store_tile<AMODE, true>(As[0], a0, a1);
  store_tile<BMODE, false>(Bs[0], b0, b1);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = k0 + BK < kend;
    if (has_next) {
      load_tile<AMODE, true>(p, m0, k0 + BK, kend, a0, a1);
      load_tile<BMODE, false>(p, n0, k0 + BK, kend, b0, b1);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 x0 = ld_f4(&As[buf][kk][ty * 4]), x1 = ld_f4(&As[buf][kk][64 + ty * 4]);
      float4 y0 = ld_f4(&Bs[buf][kk][tx * 4]), y1 = ld_f4(&Bs[buf][kk][64 + tx * 4]);
      const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float yb[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xa[i], yb[j], acc[i][j]);
    }
    if (has_next) {
      store_tile<AMODE, true>(As[buf ^ 1], a0, a1);
      store_tile<BMODE, false>(Bs[buf ^ 1], b0, b1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* C = p.C + (size_t)blockIdx.z * p.M * p.ldc;
  const bool raw = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    float rs = (!raw && p.rowscale) ? p.rowscale[m] : 1.f;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int n = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (!raw) {
          if (p.bias && n + j < p.N) x += p.bias[n + j];
          if (p.relu) x = fmaxf(x, 0.f);
          if (p.rowscale) x *= rs;
          if (p.mask_aux && n + j < p.N) x = p.mask_aux[(size_t)m * p.ld_aux + n + j] > 0.f ? x : 0.f;
        }
        v[j] = x;
      }
      float* dst = C + (size_t)m * p.ldc + n;
      if (n + 3 < p.N && (p.ldc & 3) == 0) st_f4(dst, make_float4(v[0], v[1], v[2], v[3]));
      else for (int j = 0; j < 4 && n + j < p.N; ++j) dst[j] = v[j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, long long MN, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[(size_t)z * MN + i];
  C[i] = acc;
}

// column sums (bias gradients): out[n] = sum_m X[m, n], ordered two-stage reduction.
// Block = 32 column groups (4 columns each) x 8 row lanes; every thread keeps 4 independent
// 16-byte loads in flight, the 8 row lanes are combined through shared memory in a fixed order.
constexpr int CS_TX = 32, CS_TY = 8, CS_VEC = 4;
__global__ void __launch_bounds__(CS_TX * CS_TY) colsum_partial_kernel(const float* __restrict__ X, int M, int N, int ld,
                                                                      int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[CS_TY][CS_TX * CS_VEC + 4];
  const int tx = threadIdx.x % CS_TX, ty = threadIdx.x / CS_TX;
  const int col = (blockIdx.x * CS_TX + tx) * CS_VEC;
  const int mbeg = blockIdx.y * rows_per_chunk, mend = min(M, mbeg + rows_per_chunk);
  float acc[CS_VEC] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && col + CS_VEC <= N;
  if (vec) {
    int m = mbeg + ty;
    for (; m + 3 * CS_TY < mend; m += 4 * CS_TY) {
      float4 a = ld_f4(X + (size_t)m * ld + col), b = ld_f4(X + (size_t)(m + CS_TY) * ld + col);
      float4 c = ld_f4(X + (size_t)(m + 2 * CS_TY) * ld + col), d = ld_f4(X + (size_t)(m + 3 * CS_TY) * ld + col);
      acc[0] += (a.x + b.x) + (c.x + d.x); acc[1] += (a.y + b.y) + (c.y + d.y);
      acc[2] += (a.z + b.z) + (c.z + d.z); acc[3] += (a.w + b.w) + (c.w + d.w);
    }
    for (; m < mend; m += CS_TY) {
      float4 a = ld_f4(X + (size_t)m * ld + col);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    }
  } else {
    for (int m = mbeg + ty; m < mend; m += CS_TY)
      for (int j = 0; j < CS_VEC; ++j)
        if (col + j < N) acc[j] += X[(size_t)m * ld + col + j];
  }
  for (int j = 0; j < CS_VEC; ++j) red[ty][tx * CS_VEC + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < CS_TX * CS_VEC) {
    const int n = blockIdx.x * CS_TX * CS_VEC + threadIdx.x;
    if (n < N) {
      float s = 0.f;
      for (int r = 0; r < CS_TY; ++r) s += red[r][threadIdx.x];
      partial[(size_t)blockIdx.y * N + n] = s;
    }
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += partial[(size_t)c * N + n];
  out[n] = acc;
}
// number of row chunks: enough blocks to fill the SMs ~4x, at least 32 rows per chunk
int colsum_chunks(int M, int N, int cols_per_block) {
  int col_blocks = csg_div_up(N, cols_per_block);
  int want = csg_div_up(4 * 148, col_blocks);
  int maxc = csg_div_up(M, 32);
  if (want > maxc) want = maxc;
  return want < 1 ? 1 : want;
}

template <int AMODE, int BMODE>
int launch(const GemmParams& p, int splits, cudaStream_t stream) {
  dim3 grid(csg_div_up(p.N, BN), csg_div_up(p.M, BM), splits);
  gemm_f32_kernel<AMODE, BMODE><<<grid, NT, 0, stream>>>(p);
  CSG_CHECK_LAUNCH("csg_gemm_f32");
  return 0;
}

}  // namespace

CSG_API size_t csg_gemm_f32_workspace(int M, int N, int K, int amode) {
  // split-K partials for the weight-gradient shape (A_COL): at most 64 splits
  if (amode != A_COL) return 0;
  return (size_t)64 * M * N * sizeof(float);
}

// amode / bmode: see the enums above.  Gather arguments are only read in the GATHER modes.
// bias / rowscale / mask_aux may be null.  workspace is needed for amode == A_COL (split-K).
CSG_API int csg_gemm_f32(int amode, int bmode, int M, int N, int K,
                         const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                         const float* bias, int relu, const float* rowscale,
                         const float* mask_aux, int ld_aux,
                         const float* g_obj, const float* g_pred, const int* g_sidx, const int* g_oidx,
                         int g_din, int g_dp, int g_ldp,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (M == 0 || N == 0) return 0;
  CSG_REQUIRE(M > 0 && N > 0 && K >= 0, "gemm_f32: bad sizes M=%d N=%d K=%d", M, N, K);
  if (K == 0 && !bias && !relu) {      // empty reduction (a batch without triples): the product is zero, no operand is read
    CSG_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, stream));
    return 0;
  }
  CsgProfScope prof(CSG_PROF_GEMM_F32, 2.0 * M * N * K, stream);
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.bias = bias; p.relu = relu; p.rowscale = rowscale; p.mask_aux = mask_aux; p.ld_aux = ld_aux;
  p.g.obj = g_obj; p.g.pred = g_pred; p.g.s_idx = g_sidx; p.g.o_idx = g_oidx; p.g.Din = g_din; p.g.Dp = g_dp; p.g.ldp = g_ldp;
  p.k_split = K > 0 ? K : 1;
  const bool gath = amode == A_GATHER || bmode == B_GATHER;
  if (gath) {
    CSG_REQUIRE(g_obj && g_pred && g_sidx && g_oidx, "gemm_f32: gather mode without gather sources");
    CSG_REQUIRE((g_din & 3) == 0 && (g_dp & 3) == 0 && (g_ldp & 3) == 0 && g_ldp >= g_dp,
                "gemm_f32: gather dims must be multiples of 4");
    int width = 2 * g_din + g_dp;
    CSG_REQUIRE(amode != A_GATHER || K == width, "gemm_f32: K=%d != 2*Din+Dp=%d", K, width);
    CSG_REQUIRE(bmode != B_GATHER || N == width, "gemm_f32: N=%d != 2*Din+Dp=%d", N, width);
  }
  if (amode == A_ROW) CSG_REQUIRE((lda & 3) == 0 && (K & 3) == 0, "gemm_f32: A_ROW needs lda, K %% 4 == 0");
  if (bmode == B_NK) CSG_REQUIRE((ldb & 3) == 0 && (K & 3) == 0, "gemm_f32: B_NK needs ldb, K %% 4 == 0");
  if (amode == A_COL) CSG_REQUIRE((lda & 3) == 0, "gemm_f32: A_COL needs lda %% 4 == 0");
  if (bmode == B_KN) CSG_REQUIRE((ldb & 3) == 0, "gemm_f32: B_KN needs ldb %% 4 == 0");

  int splits = 1;
  float* out = C;
  if (amode == A_COL && K > 256) {
    // weight gradients: tiny [M, N], long K -> split K so that the grid covers the 148 SMs
    int tiles = csg_div_up(M, BM) * csg_div_up(N, BN);
    splits = csg_div_up(2 * csg_num_sms(), tiles);
    if (splits > 64) splits = 64;
    int per = csg_div_up(csg_div_up(K, splits), BK) * BK;
    splits = csg_div_up(K, per);
    p.k_split = per;
    if (splits > 1) {
      CSG_REQUIRE(ldc == N, "gemm_f32: split-K output must be contiguous");
      CSG_REQUIRE(workspace && workspace_bytes >= (size_t)splits * M * N * sizeof(float), "gemm_f32: workspace too small");
      CSG_REQUIRE(!bias && !relu && !rowscale && !mask_aux, "gemm_f32: split-K has no epilogue");
      p.C = reinterpret_cast<float*>(workspace);
    }
  }
  int rc = 1;
  if (amode == A_ROW && bmode == B_NK) rc = launch<A_ROW, B_NK>(p, splits, stream);
  else if (amode == A_ROW && bmode == B_KN) rc = launch<A_ROW, B_KN>(p, splits, stream);
  else if (amode == A_GATHER && bmode == B_NK) rc = launch<A_GATHER, B_NK>(p, splits, stream);
  else if (amode == A_COL && bmode == B_KN) rc = launch<A_COL, B_KN>(p, splits, stream);
  else if (amode == A_COL && bmode == B_GATHER) rc = launch<A_COL, B_GATHER>(p, splits, stream);
  else { csg_set_error("gemm_f32: unsupported mode pair (%d, %d)", amode, bmode); return 1; }
  if (rc) return rc;
  if (splits > 1) {
    long long MN = (long long)M * N;
    splitk_reduce_kernel<<<csg_div_up(MN, 256), 256, 0, stream>>>(p.C, out, MN, splits);
    CSG_CHECK_LAUNCH("csg_gemm_f32 split-K reduce");
  }
  return 0;
}

__global__ void relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out,
                                 long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// out = dy * [y > 0]   (backward of the final ReLU of net2, layers.py:22-23)
CSG_API int csg_relu_mask_f32(const float* dy, const float* y, float* out, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  relu_mask_kernel<<<csg_div_up(n, 256), 256, 0, stream>>>(dy, y, out, n);
  CSG_CHECK_LAUNCH("csg_relu_mask_f32");
  return 0;
}

CSG_API size_t csg_colsum_f32_workspace(int M, int N) {
  return (size_t)colsum_chunks(M, N, CS_TX * CS_VEC) * N * sizeof(float) + 16;
}

CSG_API int csg_colsum_f32(const float* X, int M, int N, int ld, float* out, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream) {
  if (N == 0) return 0;
  const int chunks = colsum_chunks(M, N, CS_TX * CS_VEC);
  CSG_REQUIRE(workspace_bytes >= csg_colsum_f32_workspace(M, N), "colsum: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  const int rows_per_chunk = csg_div_up(M > 0 ? M : 1, chunks);
  colsum_partial_kernel<<<dim3(csg_div_up(N, CS_TX * CS_VEC), chunks), CS_TX * CS_TY, 0, stream>>>(X, M, N, ld, rows_per_chunk,
                                                                                              partial);
  CSG_CHECK_LAUNCH("csg_colsum partial");
  colsum_final_kernel<<<csg_div_up(N, 128), 128, 0, stream>>>(partial, chunks, N, out);
  CSG_CHECK_LAUNCH("csg_colsum final");
  return 0;
}
