// Single-pass layout compositor (boxes_to_layout / masks_to_layout) for sm_100a.
//
// Replaces sg2im/layout.py:12-188 of the reference: instead of materialising the
// per-object [O, D, H, W] samples (grid_sample) and scatter-adding them, each CTA
// owns one 8x64 pixel tile of one image's [D, H, W] canvas, builds the ordered
// list of objects whose support intersects the tile, evaluates the separable
// bilinear weights S_o(y, x) once per pixel (they do not depend on the channel),
// accumulates sum_o vec[o, d] * S_o(y, x) in registers and writes every canvas
// element exactly once with 128-bit streaming stores.
//
// Coordinate chain (kept operation-for-operation so that ramp pixels of small
// boxes agree with the reference, SURVEY.md §7 "layout coordinate fidelity"):
//   u  = (lin[x] - x0) / w            layout.py:101   (lin = torch.linspace(0,1,W), passed in)
//   g  = 2u - 1                       layout.py:110
//   ix = ((g + 1) * S - 1) / 2        ATen grid_sampler unnormalize, align_corners=False
//   ix = ((g + 1) / 2) * (S - 1)      align_corners=True (torch <= 1.2 behaviour)
// with S = 8 for boxes_to_layout (layout.py:34) and S = M for masks_to_layout.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int TILE_W = 64;
constexpr int TILE_H = 8;
constexpr int TILE_PX = TILE_W * TILE_H;
constexpr int NTHREADS = 256;

struct LayoutParams {
  const float* vecs;     // [NO, D]
  const float* boxes;    // [NO, 4] xywh
  const float* masks;    // [NO, M, M] or nullptr (boxes_to_layout)
  const int* obj_off;    // [N + 1]
  const float* lin_x;    // [W]
  const float* lin_y;    // [H]
  int N, D, H, W, M, align;
  int tiles_x, tiles_y;
  int lcap;              // object-list capacity held in shared memory
};

struct Tap {
  int i0;        // index of the first tap (second is i0 + 1); -2 when both are out of range
  float w0, w1;  // bilinear weights of the two taps (NaN for degenerate boxes, as in ATen)
};

__device__ __forceinline__ Tap axis_tap(float lin, float start, float extent, int size, int align) {
  float u = __fdiv_rn(__fsub_rn(lin, start), extent);
  float g = __fsub_rn(__fmul_rn(u, 2.f), 1.f);
  float ix;
  if (align) ix = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  else       ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
  float f = floorf(ix);
  float t = __fsub_rn(ix, f);
  Tap r;
  r.w0 = __fsub_rn(1.f, t);
  r.w1 = t;
  r.i0 = (f >= -2.f && f <= (float)size) ? (int)f : -2;
  return r;
}

// Conservative test: can an object with (start, extent) touch linspace range [lo, hi]?
// Anything not provably outside (including NaN / zero extents) is kept, so culling never
// changes a result: a kept object that does not touch a pixel contributes an exact 0.
__device__ __forceinline__ bool axis_may_touch(float start, float extent, float lo, float hi,
                                               int size, int align) {
  float m = align ? (size > 1 ? 1.f / (float)(size - 1) : INFINITY) : 0.5f / (float)size;
  float a = start - m * extent, b = start + (1.f + m) * extent;
  float mn = fminf(a, b), mx = fmaxf(a, b);
  float eps = 1e-4f * (fabsf(start) + fabsf(extent) + 1.f);
  if (!(extent > 0.f || extent < 0.f)) return true;
  if (!(fabsf(a) < INFINITY) || !(fabsf(b) < INFINITY)) return true;
  return !(mx + eps < lo || mn - eps > hi);
}

struct Smem {
  int* list;     // [lcap]            object ids (global) in ascending order
  int* ci;       // [lcap][TILE_W]    first column tap
  float* cw;     // [lcap][TILE_W][2] column tap weights
  int* ri;       // [lcap][TILE_H]
  float* rw;     // [lcap][TILE_H][2]
  float* wS;     // [lcap][TILE_PX]   S_o(y, x) for the tile
  float* vS;     // [lcap][D]         (forward: vecs; backward: per-object accumulators)
};

__device__ __forceinline__ Smem carve(float* base, int lcap, int D) {
  Smem s;
  s.wS = base;                                  // 16B aligned: first
  s.vS = s.wS + (size_t)lcap * TILE_PX;
  s.cw = s.vS + (size_t)lcap * D;
  s.rw = s.cw + (size_t)lcap * TILE_W * 2;
  s.ci = reinterpret_cast<int*>(s.rw + (size_t)lcap * TILE_H * 2);
  s.ri = s.ci + (size_t)lcap * TILE_W;
  s.list = s.ri + (size_t)lcap * TILE_H;
  return s;
}

size_t smem_bytes(int lcap, int D) {
  return sizeof(float) * ((size_t)lcap * TILE_PX + (size_t)lcap * D + (size_t)lcap * TILE_W * 2 +
                          (size_t)lcap * TILE_H * 2 + (size_t)lcap * TILE_W + (size_t)lcap * TILE_H + lcap);
}

// Warp 0 appends, in ascending object order, the objects of [*cursor, oend) that may touch
// the tile, until the list holds lcap entries.  Returns through shared memory.
__device__ void build_list(const LayoutParams& p, const Smem& s, int oend, int x0, int y0,
                           int* s_cursor, int* s_count) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int S = p.masks ? p.M : 8;
    const float xlo = p.lin_x[x0], xhi = p.lin_x[min(x0 + TILE_W, p.W) - 1];
    const float ylo = p.lin_y[y0], yhi = p.lin_y[min(y0 + TILE_H, p.H) - 1];
    int cursor = *s_cursor, count = 0;
    while (cursor < oend && count < p.lcap) {
      int o = cursor + lane;
      bool keep = false;
      if (o < oend) {
        float4 b = ld_f4(p.boxes + 4 * (size_t)o);
        // zero / NaN extents and non-finite origins give NaN weights that poison every pixel of the image
        // (0 * NaN in grid_sample), whatever the other axis says: never cull those
        const bool poison = !(b.z > 0.f || b.z < 0.f) || !(b.w > 0.f || b.w < 0.f) ||
                            !(fabsf(b.x) < INFINITY) || !(fabsf(b.y) < INFINITY);
        keep = poison ||
               (axis_may_touch(b.x, b.z, xlo, xhi, S, p.align) && axis_may_touch(b.y, b.w, ylo, yhi, S, p.align));
      }
      unsigned bal = __ballot_sync(0xffffffffu, keep);
      int pos = count + __popc(bal & ((1u << lane) - 1u));
      int room = p.lcap - count;
      int total = __popc(bal);
      if (total <= room) {
        if (keep) s.list[pos] = o;
        count += total;
        cursor += 32;
      } else {
        // take only the first `room` kept objects; resume after the last one taken
        if (keep && pos < p.lcap) s.list[pos] = o;
        unsigned taken_last = __ballot_sync(0xffffffffu, keep && pos == p.lcap - 1);
        int last_lane = __ffs(taken_last) - 1;
        cursor += last_lane + 1;
        count = p.lcap;
      }
    }
    if (lane == 0) { *s_cursor = min(cursor, oend); *s_count = count; }
  }
}

template <bool HAS_MASK>
__device__ void build_weights(const LayoutParams& p, const Smem& s, int L, int x0, int y0) {
  const int tid = threadIdx.x;
  const int S = HAS_MASK ? p.M : 8;
  // separable taps: 64 column + 8 row evaluations per object
  for (int i = tid; i < L * (TILE_W + TILE_H); i += NTHREADS) {
    int c = i / (TILE_W + TILE_H), r = i % (TILE_W + TILE_H);
    float4 b = ld_f4(p.boxes + 4 * (size_t)s.list[c]);
    if (r < TILE_W) {
      int x = x0 + r;
      Tap t = axis_tap(p.lin_x[min(x, p.W - 1)], b.x, b.z, S, p.align);
      s.ci[c * TILE_W + r] = t.i0;
      s.cw[(c * TILE_W + r) * 2] = t.w0;
      s.cw[(c * TILE_W + r) * 2 + 1] = t.w1;
    } else {
      r -= TILE_W;
      int y = y0 + r;
      Tap t = axis_tap(p.lin_y[min(y, p.H - 1)], b.y, b.w, S, p.align);
      s.ri[c * TILE_H + r] = t.i0;
      s.rw[(c * TILE_H + r) * 2] = t.w0;
      s.rw[(c * TILE_H + r) * 2 + 1] = t.w1;
    }
  }
  __syncthreads();
  for (int i = tid; i < L * TILE_PX; i += NTHREADS) {
    int c = i / TILE_PX, px = i % TILE_PX;
    int row = px / TILE_W, col = px % TILE_W;
    int ix = s.ci[c * TILE_W + col], iy = s.ri[c * TILE_H + row];
    float wx0 = s.cw[(c * TILE_W + col) * 2], wx1 = s.cw[(c * TILE_W + col) * 2 + 1];
    float wy0 = s.rw[(c * TILE_H + row) * 2], wy1 = s.rw[(c * TILE_H + row) * 2 + 1];
    bool vx0 = ix >= 0 && ix < S, vx1 = ix >= -1 && ix < S - 1;
    bool vy0 = iy >= 0 && iy < S, vy1 = iy >= -1 && iy < S - 1;
    float w;
    if (HAS_MASK) {
      const float* m = p.masks + (size_t)s.list[c] * S * S;
      float m00 = (vy0 && vx0) ? __ldg(m + iy * S + ix) : 0.f;
      float m01 = (vy0 && vx1) ? __ldg(m + iy * S + ix + 1) : 0.f;
      float m10 = (vy1 && vx0) ? __ldg(m + (iy + 1) * S + ix) : 0.f;
      float m11 = (vy1 && vx1) ? __ldg(m + (iy + 1) * S + ix + 1) : 0.f;
      // nw, ne, sw, se order of ATen's bilinear
      w = m00 * (wx0 * wy0) + m01 * (wx1 * wy0) + m10 * (wx0 * wy1) + m11 * (wx1 * wy1);
    } else {
      float ax = (vx0 ? 1.f : 0.f) * wx0 + (vx1 ? 1.f : 0.f) * wx1;
      float ay = (vy0 ? 1.f : 0.f) * wy0 + (vy1 ? 1.f : 0.f) * wy1;
      w = ax * ay;
    }
    s.wS[i] = w;
  }
}

// ----------------------------------------------------------------------------------------
// forward: out[n, d, y, x] = sum_o vecs[o, d] * S_o(y, x)
// ----------------------------------------------------------------------------------------
template <bool HAS_MASK>
__global__ void __launch_bounds__(NTHREADS) layout_fwd_kernel(LayoutParams p, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ int s_cursor, s_count;
  const Smem s = carve(smem_raw, p.lcap, p.D);
  const int n = blockIdx.y;
  const int x0 = (blockIdx.x % p.tiles_x) * TILE_W, y0 = (blockIdx.x / p.tiles_x) * TILE_H;
  const int oend = p.obj_off[n + 1];
  const int tid = threadIdx.x;
  if (tid == 0) s_cursor = p.obj_off[n];
  __syncthreads();

  const int q = tid & 127, half = tid >> 7;
  const int row = q >> 4, col4 = (q & 15) * 4;
  const int y = y0 + row, x = x0 + col4;
  const bool vec4 = (p.W & 3) == 0;
  const bool active = y < p.H && x < p.W;
  bool first = true;
  while (true) {
    build_list(p, s, oend, x0, y0, &s_cursor, &s_count);
    __syncthreads();
    const int L = s_count;
    const bool more = s_cursor < oend;   // read before the next build_list may advance it
    build_weights<HAS_MASK>(p, s, L, x0, y0);
    for (int i = tid; i < L * p.D; i += NTHREADS) {
      int c = i / p.D, d = i % p.D;
      s.vS[i] = p.vecs[(size_t)s.list[c] * p.D + d];
    }
    __syncthreads();
    if (active) {
      for (int dg = half * 4; dg < p.D; dg += 8) {
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
        const int nd = min(4, p.D - dg);
        if (nd == 4) {
          for (int c = 0; c < L; ++c) {
            float4 w = ld_f4(s.wS + c * TILE_PX + q * 4);
            float4 v = ld_f4(s.vS + c * p.D + dg);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j][0] = fmaf(vv[j], w.x, acc[j][0]);
              acc[j][1] = fmaf(vv[j], w.y, acc[j][1]);
              acc[j][2] = fmaf(vv[j], w.z, acc[j][2]);
              acc[j][3] = fmaf(vv[j], w.w, acc[j][3]);
            }
          }
        } else {
          for (int c = 0; c < L; ++c) {
            float4 w = ld_f4(s.wS + c * TILE_PX + q * 4);
            for (int j = 0; j < nd; ++j) {
              float vj = s.vS[c * p.D + dg + j];
              acc[j][0] = fmaf(vj, w.x, acc[j][0]);
              acc[j][1] = fmaf(vj, w.y, acc[j][1]);
              acc[j][2] = fmaf(vj, w.z, acc[j][2]);
              acc[j][3] = fmaf(vj, w.w, acc[j][3]);
            }
          }
        }
        for (int j = 0; j < nd; ++j) {
          float* dst = out + (((size_t)n * p.D + dg + j) * p.H + y) * p.W + x;
          if (vec4) {
            float4 r = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            if (!first) { float4 o = ld_f4(dst); r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
            st_f4_stream(dst, r);
          } else {
            for (int k = 0; k < 4 && x + k < p.W; ++k) dst[k] = first ? acc[j][k] : dst[k] + acc[j][k];
          }
        }
      }
    }
    first = false;
    __syncthreads();
    if (!more) break;
  }
}

// ----------------------------------------------------------------------------------------
// backward wrt vecs: dvecs[o, d] = sum_{y, x} dout[n, d, y, x] * S_o(y, x)
// Pass 1 (per tile): partial[tile][o_local][d]; pass 2: ordered sum over the tiles of an image.
// ----------------------------------------------------------------------------------------
constexpr int BWD_CH = 8;   // objects whose 8 weights per pixel slice are kept in registers

template <bool HAS_MASK>
__global__ void __launch_bounds__(NTHREADS) layout_bwd_vecs_kernel(LayoutParams p, const float* __restrict__ dout,
                                                                  float* __restrict__ partial) {
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ int s_cursor, s_count;
  const Smem s = carve(smem_raw, p.lcap, p.D);
  const int n = blockIdx.y;
  const int x0 = (blockIdx.x % p.tiles_x) * TILE_W, y0 = (blockIdx.x / p.tiles_x) * TILE_H;
  const int obeg = p.obj_off[n], oend = p.obj_off[n + 1];
  const int On = oend - obeg;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  float* my_partial = partial + ((size_t)tiles * obeg + (size_t)blockIdx.x * On) * p.D;
  for (int i = tid; i < On * p.D; i += NTHREADS) my_partial[i] = 0.f;
  if (tid == 0) s_cursor = obeg;
  __syncthreads();

  const int slice = tid & 63, cg = tid >> 6;       // 64 slices of 8 px, 4 channel groups
  const int row = slice >> 3, col8 = (slice & 7) * 8;
  const int y = y0 + row, x = x0 + col8;
  const bool vec4 = (p.W & 3) == 0;
  const bool rowok = y < p.H;
  while (true) {
    build_list(p, s, oend, x0, y0, &s_cursor, &s_count);
    __syncthreads();
    const int L = s_count;
    const bool more = s_cursor < oend;
    build_weights<HAS_MASK>(p, s, L, x0, y0);
    for (int i = tid; i < L * p.D; i += NTHREADS) s.vS[i] = 0.f;
    __syncthreads();
    for (int cb = 0; cb < L; cb += BWD_CH) {
      float wr[BWD_CH][8];
#pragma unroll
      for (int c = 0; c < BWD_CH; ++c) {
        if (cb + c < L) {
          float4 a = ld_f4(s.wS + (cb + c) * TILE_PX + slice * 8);
          float4 b = ld_f4(s.wS + (cb + c) * TILE_PX + slice * 8 + 4);
          wr[c][0] = a.x; wr[c][1] = a.y; wr[c][2] = a.z; wr[c][3] = a.w;
          wr[c][4] = b.x; wr[c][5] = b.y; wr[c][6] = b.z; wr[c][7] = b.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) wr[c][k] = 0.f;
        }
      }
      for (int d = cg; d < p.D; d += 4) {
        float g[8];
        const float* src = dout + (((size_t)n * p.D + d) * p.H + y) * p.W + x;
        if (rowok && vec4 && x + 7 < p.W) {
          float4 a = ld_f4_stream(src), b = ld_f4_stream(src + 4);
          g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] = (rowok && x + k < p.W) ? src[k] : 0.f;
        }
        float v[BWD_CH];
#pragma unroll
        for (int c = 0; c < BWD_CH; ++c) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) a = fmaf(g[k], wr[c][k], a);
          v[c] = a;
        }
        // transposing butterfly: 8 values x 32 lanes -> lane holds the warp sum of object
        // c = 4*bit4 + 2*bit3 + bit2 of its lane id (9 shuffles instead of 40)
        {
          const bool hi = lane & 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float send = hi ? v[i] : v[i + 4];
            float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            v[i] = (hi ? v[i + 4] : v[i]) + recv;
          }
        }
        {
          const bool hi = lane & 8;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float send = hi ? v[i] : v[i + 2];
            float recv = __shfl_xor_sync(0xffffffffu, send, 8);
            v[i] = (hi ? v[i + 2] : v[i]) + recv;
          }
        }
        {
          const bool hi = lane & 4;
          float send = hi ? v[0] : v[1];
          float recv = __shfl_xor_sync(0xffffffffu, send, 4);
          v[0] = (hi ? v[1] : v[0]) + recv;
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        if ((lane & 3) == 0) {
          int c = cb + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          // two warps (the two 256-pixel halves of the tile) add into each cell: a + b is order independent
          if (c < L) atomicAdd(&s.vS[c * p.D + d], v[0]);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < L * p.D; i += NTHREADS) {
      int c = i / p.D, d = i % p.D;
      my_partial[(size_t)(s.list[c] - obeg) * p.D + d] = s.vS[i];
    }
    __syncthreads();
    if (!more) break;
  }
}

__global__ void layout_bwd_reduce_kernel(const float* __restrict__ partial, const int* __restrict__ obj_off,
                                         const int* __restrict__ obj_img, float* __restrict__ dvecs,
                                         int NO, int D, int tiles) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)NO * D) return;
  int o = (int)(i / D), d = (int)(i % D);
  int n = obj_img[o];
  int obeg = obj_off[n], On = obj_off[n + 1] - obeg;
  const float* src = partial + ((size_t)tiles * obeg) * D + (size_t)(o - obeg) * D + d;
  float acc = 0.f;
  for (int t = 0; t < tiles; ++t) acc += src[(size_t)t * On * D];
  dvecs[i] = acc;
}

__global__ void obj_img_kernel(const int* __restrict__ obj_off, int N, int* __restrict__ obj_img) {
  int n = blockIdx.x;
  for (int o = obj_off[n] + threadIdx.x; o < obj_off[n + 1]; o += blockDim.x) obj_img[o] = n;
}

int pick_lcap(int max_objs, int D, size_t extra = 0) {
  int lcap = max_objs > 0 ? max_objs : 16;
  if (lcap < 4) lcap = 4;
  if (lcap > 48) lcap = 48;
  while (lcap > 4 && smem_bytes(lcap, D) + extra > 200 * 1024) lcap -= 4;
  return lcap;
}

int fill_params(LayoutParams& p, const float* vecs, const float* boxes, const float* masks, const int* obj_off,
                const float* lin_x, const float* lin_y, int N, int D, int H, int W, int M, int align,
                int max_objs) {
  CSG_REQUIRE(N >= 0 && D > 0 && H > 0 && W > 0, "layout: bad sizes N=%d D=%d H=%d W=%d", N, D, H, W);
  CSG_REQUIRE(masks == nullptr || M > 0, "layout: masks given but M=%d", M);
  CSG_REQUIRE((D & 3) == 0, "layout: D=%d must be a multiple of 4", D);
  p.vecs = vecs; p.boxes = boxes; p.masks = masks; p.obj_off = obj_off; p.lin_x = lin_x; p.lin_y = lin_y;
  p.N = N; p.D = D; p.H = H; p.W = W; p.M = M; p.align = align;
  p.tiles_x = csg_div_up(W, TILE_W); p.tiles_y = csg_div_up(H, TILE_H);
  p.lcap = pick_lcap(max_objs, D);
  return 0;
}

}  // namespace

// ----------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------
CSG_API int csg_layout_fwd(const float* vecs, const float* boxes, const float* masks, const int* obj_off,
                           const float* lin_x, const float* lin_y, float* out, int N, int D, int H, int W,
                           int M, int align_corners, int max_objs_per_image, cudaStream_t stream) {
  LayoutParams p;
  if (int rc = fill_params(p, vecs, boxes, masks, obj_off, lin_x, lin_y, N, D, H, W, M, align_corners,
                           max_objs_per_image)) return rc;
  if (N == 0) return 0;
  size_t smem = smem_bytes(p.lcap, D);
  dim3 grid(p.tiles_x * p.tiles_y, N);
  if (masks) {
    CSG_CUDA(cudaFuncSetAttribute(layout_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_fwd_kernel<true><<<grid, NTHREADS, smem, stream>>>(p, out);
  } else {
    CSG_CUDA(cudaFuncSetAttribute(layout_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_fwd_kernel<false><<<grid, NTHREADS, smem, stream>>>(p, out);
  }
  CSG_CHECK_LAUNCH("csg_layout_fwd");
  return 0;
}

CSG_API size_t csg_layout_bwd_vecs_workspace(int NO, int D, int H, int W) {
  size_t tiles = (size_t)csg_div_up(W, TILE_W) * csg_div_up(H, TILE_H);
  return tiles * (size_t)NO * D * sizeof(float) + (size_t)(NO + 1) * sizeof(int) + 256;
}

CSG_API int csg_layout_bwd_vecs(const float* dout, const float* boxes, const float* masks, const int* obj_off,
                                const float* lin_x, const float* lin_y, float* dvecs, int N, int NO, int D,
                                int H, int W, int M, int align_corners, int max_objs_per_image,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LayoutParams p;
  if (int rc = fill_params(p, nullptr, boxes, masks, obj_off, lin_x, lin_y, N, D, H, W, M, align_corners,
                           max_objs_per_image)) return rc;
  if (N == 0 || NO == 0) return 0;
  CSG_REQUIRE(workspace_bytes >= csg_layout_bwd_vecs_workspace(NO, D, H, W), "layout bwd: workspace too small");
  const int tiles = p.tiles_x * p.tiles_y;
  float* partial = reinterpret_cast<float*>(workspace);
  int* obj_img = reinterpret_cast<int*>(partial + (size_t)tiles * NO * D);
  size_t smem = smem_bytes(p.lcap, D);
  dim3 grid(tiles, N);
  obj_img_kernel<<<N, 64, 0, stream>>>(obj_off, N, obj_img);
  if (masks) {
    CSG_CUDA(cudaFuncSetAttribute(layout_bwd_vecs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_bwd_vecs_kernel<true><<<grid, NTHREADS, smem, stream>>>(p, dout, partial);
  } else {
    CSG_CUDA(cudaFuncSetAttribute(layout_bwd_vecs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_bwd_vecs_kernel<false><<<grid, NTHREADS, smem, stream>>>(p, dout, partial);
  }
  CSG_CHECK_LAUNCH("csg_layout_bwd_vecs");
  layout_bwd_reduce_kernel<<<csg_div_up((long long)NO * D, 256), 256, 0, stream>>>(partial, obj_off, obj_img, dvecs,
                                                                                  NO, D, tiles);
  CSG_CHECK_LAUNCH("csg_layout_bwd_reduce");
  return 0;
}
