// Coordinate chain and bilinear taps shared by the layout compositor kernels (layout.cu, layout_geom.cu).
#pragma once
#include "common.cuh"
#include <math.h>

namespace {

struct LayoutParams {
  const float* vecs;     // [NO, D]
  const float* boxes;    // [NO, 4] xywh
  const float* masks;    // [NO, M, M] or nullptr (boxes_to_layout)
  const int* obj_off;    // [N + 1]
  const float* lin_x;    // [W]
  const float* lin_y;    // [H]
  int N, D, H, W, M, align;
  int TW, TH;            // tile width / height (forward: 64 x 8 or 128 x 4; generic backward: 64 x 8)
  int tiles_x, tiles_y;
  int lcap;              // object-list capacity held in shared memory
};

// unnormalised sample coordinate of ATen's grid_sampler for one axis
__device__ __forceinline__ float axis_coord(float lin, float start, float extent, int size, int align) {
  float u = __fdiv_rn(__fsub_rn(lin, start), extent);
  float g = __fsub_rn(__fmul_rn(u, 2.f), 1.f);
  if (align) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
}

struct Tap {
  int i0;        // index of the first tap (second is i0 + 1); -2 when both are out of range
  float w0, w1;  // bilinear weights of the two taps (NaN for degenerate boxes, as in ATen)
};

__device__ __forceinline__ Tap coord_tap(float ix, int size) {
  float f = floorf(ix);
  float t = __fsub_rn(ix, f);
  Tap r;
  r.w0 = __fsub_rn(1.f, t);
  r.w1 = t;
  r.i0 = (f >= -2.f && f <= (float)size) ? (int)f : -2;
  return r;
}

__device__ __forceinline__ Tap axis_tap(float lin, float start, float extent, int size, int align) {
  return coord_tap(axis_coord(lin, start, extent, size, align), size);
}

// weight of a constant-1 source sampled with zero padding (boxes_to_layout): sum of the in-range tap weights
__device__ __forceinline__ float tap_ones(const Tap& t, int size) {
  const bool v0 = t.i0 >= 0 && t.i0 < size, v1 = t.i0 >= -1 && t.i0 < size - 1;
  return (v0 ? 1.f : 0.f) * t.w0 + (v1 ? 1.f : 0.f) * t.w1;
}

// 4-tap bilinear read of an S x S mask with zero padding, in ATen's nw, ne, sw, se order
__device__ __forceinline__ float mask_weight(const float* __restrict__ m, int S, const Tap& tx, const Tap& ty) {
  const int ix = tx.i0, iy = ty.i0;
  const bool vx0 = ix >= 0 && ix < S, vx1 = ix >= -1 && ix < S - 1;
  const bool vy0 = iy >= 0 && iy < S, vy1 = iy >= -1 && iy < S - 1;
  float m00 = (vy0 && vx0) ? __ldg(m + iy * S + ix) : 0.f;
  float m01 = (vy0 && vx1) ? __ldg(m + iy * S + ix + 1) : 0.f;
  float m10 = (vy1 && vx0) ? __ldg(m + (iy + 1) * S + ix) : 0.f;
  float m11 = (vy1 && vx1) ? __ldg(m + (iy + 1) * S + ix + 1) : 0.f;
  return m00 * (tx.w0 * ty.w0) + m01 * (tx.w1 * ty.w0) + m10 * (tx.w0 * ty.w1) + m11 * (tx.w1 * ty.w1);
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

__device__ __forceinline__ bool box_poison(float4 b) {
  // zero / NaN extents and non-finite origins give NaN weights that poison every pixel of the image
  // (0 * NaN in grid_sample), whatever the other axis says: such objects are never culled or skipped
  return !(b.z > 0.f || b.z < 0.f) || !(b.w > 0.f || b.w < 0.f) || !(fabsf(b.x) < INFINITY) || !(fabsf(b.y) < INFINITY);
}

}  // namespace
