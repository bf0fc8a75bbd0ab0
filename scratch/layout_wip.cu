// Single-pass layout compositor (boxes_to_layout / masks_to_layout) for sm_100a.
//
// Replaces sg2im/layout.py:12-188 of the reference: instead of materialising the per-object
// [O, D, H, W] samples (grid_sample) and scatter-adding them, a CTA owns a tile of one image's
// [D, H, W] canvas, builds the ordered list of objects whose support intersects the tile, and
//
//   forward   4 warps, each owning a 128-pixel strip (2 rows x 64 or 1 row x 128): the warp compacts the tile
//             list to the objects that touch ITS strip, every lane accumulates 4 pixels x 16 channels in
//             registers (64 FMAs per 1 weight + 4 broadcast vector loads from shared memory) and writes each
//             canvas element exactly once with 128-bit streaming stores;
//   backward  (d/dvecs) one thread per channel: the incoming gradient is streamed once with 128-bit loads, the
//             pixel weights are warp-uniform, so objects that do not touch a 4-pixel group are skipped without
//             divergence; per-tile partial sums are combined across tiles in a fixed order.
//
// The pixel weight of object o is separable for boxes_to_layout, S_o(y, x) = ay_o(y) * ax_o(x), and a 4-tap
// bilinear read of the mask for masks_to_layout (kept per tile in shared memory).
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

constexpr int NTHREADS = 128;          // forward: 4 warps x 128-pixel strips; backward: one thread per channel slot
constexpr int STRIP_PX = 128;
constexpr int TILE_PX = 4 * STRIP_PX;  // 512 pixels per tile: 8 rows x 64 or 4 rows x 128

struct LayoutParams {
  const float* vecs;     // [NO, D]
  const float* boxes;    // [NO, 4] xywh
  const float* masks;    // [NO, M, M] or nullptr (boxes_to_layout)
  const int* obj_off;    // [N + 1]
  const float* lin_x;    // [W]
  const float* lin_y;    // [H]
  int N, D, H, W, M, align;
  int TW, TH;            // tile: 64 x 8 (W <= 64) or 128 x 4
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
  float* wS;     // [lcap][TILE_PX]   S_o(y, x) for the tile (masks_to_layout only)
  float* vS;     // [lcap][D]         forward: vecs; backward: per-object accumulators
  float* ax;     // [lcap][TW]        column factor (boxes) / unused (masks)
  float* ay;     // [lcap][TH]        row factor
  int* list;     // [lcap]            object ids (global) in ascending order
  int* flags;    // [lcap]            bit 0: poison (never skipped); bits 8..: 4-pixel column groups touched
  unsigned* rowmask;   // [lcap]      rows of the tile touched
  unsigned char* act;  // [4][lcap]   per-warp compacted list (forward)
  int* nact;     // [4]
};

__host__ __device__ inline size_t smem_floats(int lcap, int D, int TW, int TH, bool mask) {
  return (size_t)lcap * ((mask ? TILE_PX : 0) + D + TW + TH + 3) + 4 + (size_t)lcap;   // act: 4*lcap bytes = lcap words
}

__device__ __forceinline__ Smem carve(float* base, int lcap, int D, int TW, int TH, bool mask) {
  Smem s;
  s.wS = base;                                  // 16B aligned: first
  s.vS = s.wS + (mask ? (size_t)lcap * TILE_PX : 0);
  s.ax = s.vS + (size_t)lcap * D;
  s.ay = s.ax + (size_t)lcap * TW;
  s.list = reinterpret_cast<int*>(s.ay + (size_t)lcap * TH);
  s.flags = s.list + lcap;
  s.rowmask = reinterpret_cast<unsigned*>(s.flags + lcap);
  s.nact = reinterpret_cast<int*>(s.rowmask + lcap);
  s.act = reinterpret_cast<unsigned char*>(s.nact + 4);
  return s;
}

__device__ __forceinline__ bool box_poison(float4 b) {
  // zero / NaN extents and non-finite origins give NaN weights that poison every pixel of the image
  // (0 * NaN in grid_sample), whatever the other axis says
  return !(b.z > 0.f || b.z < 0.f) || !(b.w > 0.f || b.w < 0.f) || !(fabsf(b.x) < INFINITY) || !(fabsf(b.y) < INFINITY);
}

// Warp 0 appends, in ascending object order, the objects of [*cursor, oend) that may touch
// the tile, until the list holds lcap entries.  Returns through shared memory.
__device__ void build_list(const LayoutParams& p, const Smem& s, int oend, int x0, int y0,
                           int* s_cursor, int* s_count) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int S = p.masks ? p.M : 8;
    const float xlo = p.lin_x[x0], xhi = p.lin_x[min(x0 + p.TW, p.W) - 1];
    const float ylo = p.lin_y[y0], yhi = p.lin_y[min(y0 + p.TH, p.H) - 1];
    int cursor = *s_cursor, count = 0;
    while (cursor < oend && count < p.lcap) {
      int o = cursor + lane;
      bool keep = false;
      if (o < oend) {
        float4 b = ld_f4(p.boxes + 4 * (size_t)o);
        keep = box_poison(b) ||
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

// Per-object separable factors for the tile (and, with masks, the per-pixel weights), plus the masks of
// rows / 4-pixel column groups where the weight can be non-zero (anything not exactly 0, NaN included).
template <bool HAS_MASK>
__device__ void build_weights(const LayoutParams& p, const Smem& s, int L, int x0, int y0) {
  const int tid = threadIdx.x;
  const int S = HAS_MASK ? p.M : 8;
  const int TW = p.TW, TH = p.TH;
  for (int c = tid; c < L; c += NTHREADS) {
    s.flags[c] = box_poison(ld_f4(p.boxes + 4 * (size_t)s.list[c])) ? 1 : 0;
    s.rowmask[c] = 0u;
  }
  __syncthreads();
  if (!HAS_MASK) {
    for (int i = tid; i < L * (TW + TH); i += NTHREADS) {
      int c = i / (TW + TH), r = i % (TW + TH);
      float4 b = ld_f4(p.boxes + 4 * (size_t)s.list[c]);
      const bool isx = r < TW;
      if (!isx) r -= TW;
      const int pos = (isx ? x0 : y0) + r;
      const int lim = isx ? p.W : p.H;
      Tap t = axis_tap((isx ? p.lin_x : p.lin_y)[min(pos, lim - 1)], isx ? b.x : b.y, isx ? b.z : b.w, S, p.align);
      const bool v0 = t.i0 >= 0 && t.i0 < S, v1 = t.i0 >= -1 && t.i0 < S - 1;
      float a = (v0 ? 1.f : 0.f) * t.w0 + (v1 ? 1.f : 0.f) * t.w1;
      if (pos >= lim) a = 0.f;
      if (isx) {
        s.ax[c * TW + r] = a;
        if (!(a == 0.f)) atomicOr(&s.flags[c], 256 << (r >> 2));
      } else {
        s.ay[c * TH + r] = a;
        if (!(a == 0.f)) atomicOr(&s.rowmask[c], 1u << r);
      }
    }
  } else {
    // separable taps first (ax / ay hold nothing in this mode; reuse them as scratch for the tap weights)
    for (int i = tid; i < L * TILE_PX; i += NTHREADS) {
      const int c = i / TILE_PX, px = i % TILE_PX;
      const int row = px / TW, col = px % TW;
      const int x = x0 + col, y = y0 + row;
      float w = 0.f;
      if (x < p.W && y < p.H) {
        float4 b = ld_f4(p.boxes + 4 * (size_t)s.list[c]);
        Tap tx = axis_tap(p.lin_x[x], b.x, b.z, S, p.align);
        Tap ty = axis_tap(p.lin_y[y], b.y, b.w, S, p.align);
        const int ix = tx.i0, iy = ty.i0;
        const bool vx0 = ix >= 0 && ix < S, vx1 = ix >= -1 && ix < S - 1;
        const bool vy0 = iy >= 0 && iy < S, vy1 = iy >= -1 && iy < S - 1;
        const float* m = p.masks + (size_t)s.list[c] * S * S;
        float m00 = (vy0 && vx0) ? __ldg(m + iy * S + ix) : 0.f;
        float m01 = (vy0 && vx1) ? __ldg(m + iy * S + ix + 1) : 0.f;
        float m10 = (vy1 && vx0) ? __ldg(m + (iy + 1) * S + ix) : 0.f;
        float m11 = (vy1 && vx1) ? __ldg(m + (iy + 1) * S + ix + 1) : 0.f;
        // nw, ne, sw, se order of ATen's bilinear
        w = m00 * (tx.w0 * ty.w0) + m01 * (tx.w1 * ty.w0) + m10 * (tx.w0 * ty.w1) + m11 * (tx.w1 * ty.w1);
      }
      s.wS[i] = w;
      if (!(w == 0.f)) {
        atomicOr(&s.flags[c], 256 << (col >> 2));
        atomicOr(&s.rowmask[c], 1u << row);
      }
    }
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------------------
// forward: out[n, d, y, x] = sum_o vecs[o, d] * S_o(y, x)
// ----------------------------------------------------------------------------------------
template <bool HAS_MASK, int CH>
__global__ void __launch_bounds__(NTHREADS) layout_fwd_kernel(LayoutParams p, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ int s_cursor, s_count;
  const Smem s = carve(smem_raw, p.lcap, p.D, p.TW, p.TH, HAS_MASK);
  const int n = blockIdx.y;
  const int TW = p.TW, TH = p.TH;
  const int x0 = (blockIdx.x % p.tiles_x) * TW, y0 = (blockIdx.x / p.tiles_x) * TH;
  const int oend = p.obj_off[n + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_cursor = p.obj_off[n];
  __syncthreads();

  // this lane's 4 pixels: strip `warp` = rows [warp*SR, warp*SR + SR) of the tile, SR = 128 / TW
  const int SR = STRIP_PX / TW;
  const int spx = lane * 4;                         // pixel offset inside the strip
  const int row = warp * SR + spx / TW, col = spx % TW;
  const int y = y0 + row, x = x0 + col;
  const bool vec4 = (p.W & 3) == 0;
  const bool active = y < p.H && x < p.W;
  const unsigned my_rows = ((1u << SR) - 1u) << (warp * SR);
  unsigned char* myact = s.act + warp * p.lcap;
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
    // per-warp compaction: objects that touch this warp's strip (poisoned objects always do)
    int na = 0;
    for (int c0 = 0; c0 < L; c0 += 32) {
      const int c = c0 + lane;
      const bool keep = c < L && ((s.rowmask[c] & my_rows) != 0u || (s.flags[c] & 1));
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) myact[na + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)c;
      na += __popc(bal);
    }
    __syncthreads();
    if (active) {
      for (int d0 = 0; d0 < p.D; d0 += CH) {
        float acc[CH][4];
#pragma unroll
        for (int j = 0; j < CH; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
        for (int a = 0; a < na; ++a) {
          const int c = myact[a];
          float4 w;
          if (HAS_MASK) {
            w = ld_f4(s.wS + c * TILE_PX + warp * STRIP_PX + spx);
          } else {
            const float wy = s.ay[c * TH + row];
            w = ld_f4(s.ax + c * TW + col);
            w.x *= wy; w.y *= wy; w.z *= wy; w.w *= wy;
          }
          const float* vp = s.vS + c * p.D + d0;
#pragma unroll
          for (int j4 = 0; j4 < CH; j4 += 4) {
            const float4 v = ld_f4(vp + j4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j4 + j][0] = fmaf(vv[j], w.x, acc[j4 + j][0]);
              acc[j4 + j][1] = fmaf(vv[j], w.y, acc[j4 + j][1]);
              acc[j4 + j][2] = fmaf(vv[j], w.z, acc[j4 + j][2]);
              acc[j4 + j][3] = fmaf(vv[j], w.w, acc[j4 + j][3]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          float* dst = out + (((size_t)n * p.D + d0 + j) * p.H + y) * p.W + x;
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
// Thread = channel d (blockIdx.z selects a group of 128 channels): the 4-pixel weights are the same for
// every lane, so the per-object skip tests are warp-uniform.
// ----------------------------------------------------------------------------------------
constexpr int BWD_G = 8;    // 4-pixel groups (128-bit loads) in flight per thread

template <bool HAS_MASK>
__global__ void __launch_bounds__(NTHREADS) layout_bwd_vecs_kernel(LayoutParams p, const float* __restrict__ dout,
                                                                  float* __restrict__ partial) {
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ int s_cursor, s_count;
  const Smem s = carve(smem_raw, p.lcap, NTHREADS, p.TW, p.TH, HAS_MASK);
  const int n = blockIdx.y;
  const int TW = p.TW, TH = p.TH;
  const int x0 = (blockIdx.x % p.tiles_x) * TW, y0 = (blockIdx.x / p.tiles_x) * TH;
  const int obeg = p.obj_off[n], oend = p.obj_off[n + 1];
  const int On = oend - obeg;
  const int tid = threadIdx.x;
  const int d = blockIdx.z * NTHREADS + tid;
  const bool dok = d < p.D;
  const int tiles = p.tiles_x * p.tiles_y;
  float* my_partial = partial + ((size_t)tiles * obeg + (size_t)blockIdx.x * On) * p.D;
  if (dok) for (int o = 0; o < On; ++o) my_partial[(size_t)o * p.D + d] = 0.f;
  if (tid == 0) s_cursor = obeg;
  __syncthreads();

  const bool vec4 = (p.W & 3) == 0;
  const int gpr = TW / 4;                       // 4-pixel groups per tile row
  const int ngroups = TH * gpr;
  const float* gbase = dout + ((size_t)n * p.D + (dok ? d : 0)) * p.H * p.W;
  while (true) {
    build_list(p, s, oend, x0, y0, &s_cursor, &s_count);
    __syncthreads();
    const int L = s_count;
    const bool more = s_cursor < oend;
    build_weights<HAS_MASK>(p, s, L, x0, y0);
    for (int c = 0; c < L; ++c) s.vS[c * NTHREADS + tid] = 0.f;     // thread-private accumulator column
    for (int g0 = 0; g0 < ngroups; g0 += BWD_G) {
      float4 g[BWD_G];
#pragma unroll
      for (int u = 0; u < BWD_G; ++u) {
        const int gi = g0 + u;
        const int row = gi / gpr, col = (gi % gpr) * 4;
        const int y = y0 + row, x = x0 + col;
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dok && gi < ngroups && y < p.H && x < p.W) {
          const float* src = gbase + (size_t)y * p.W + x;
          if (vec4) g[u] = ld_f4_stream(src);
          else {
            g[u].x = src[0];
            if (x + 1 < p.W) g[u].y = src[1];
            if (x + 2 < p.W) g[u].z = src[2];
            if (x + 3 < p.W) g[u].w = src[3];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < BWD_G; ++u) {
        const int gi = g0 + u;
        if (gi >= ngroups) break;
        const int row = gi / gpr, cg = gi % gpr;
        for (int c = 0; c < L; ++c) {
          const int fl = s.flags[c];
          // warp-uniform skip: exact zeros only (poisoned objects are never skipped)
          if (!(fl & 1) && !(((s.rowmask[c] >> row) & 1u) && ((fl >> (8 + cg)) & 1))) continue;
          float4 w;
          if (HAS_MASK) {
            w = ld_f4(s.wS + c * TILE_PX + row * TW + cg * 4);
          } else {
            const float wy = s.ay[c * TH + row];
            w = ld_f4(s.ax + c * TW + cg * 4);
            w.x *= wy; w.y *= wy; w.z *= wy; w.w *= wy;
          }
          float a = s.vS[c * NTHREADS + tid];
          a = fmaf(g[u].x, w.x, a); a = fmaf(g[u].y, w.y, a); a = fmaf(g[u].z, w.z, a); a = fmaf(g[u].w, w.w, a);
          s.vS[c * NTHREADS + tid] = a;
        }
      }
    }
    if (dok) for (int c = 0; c < L; ++c) my_partial[(size_t)(s.list[c] - obeg) * p.D + d] = s.vS[c * NTHREADS + tid];
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
