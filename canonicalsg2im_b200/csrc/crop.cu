// Differentiable box crops: sg2im/bilinear.py:65-94 (crop_bbox, backend='cudnn'), :44-62 (crop_bbox_batch_cudnn),
// :155-184 (tensor_linspace).
//
// The reference expands every image once per object ([sum O, C, H, W] copies, bilinear.py:54) and calls
// F.grid_sample; here crops are ragged over images (crop_off[n]..crop_off[n+1] sample image n) and read the image
// in place.  The backward pass wrt the images is a deterministic gather: a thread owns one image pixel and walks
// the crops of its image, so no atomics are needed.
//
// Coordinate chain, kept operation for operation:
//   p0 = 2*x0 - 1, p1 = 2*(x0 + w) - 1            metrics.py:4-8, bilinear.py:84
//   X[k] = sw[k]*p0 + ew[k]*p1                     tensor_linspace: sw = linspace(1,0,WW), ew = linspace(0,1,WW) (passed in)
//   ix = ((X + 1) * W - 1) / 2                     ATen grid_sampler unnormalize (align_corners=False), or ((X+1)/2)*(W-1)
#include "common.cuh"
#include <math.h>

namespace {

struct CropParams {
  const float* feats;    // [N, C, H, W]
  const float* bbox;     // [NC, 4] xywh
  const int* crop_off;   // [N + 1]
  const float *swx, *ewx, *swy, *ewy;   // [WW], [WW], [HH], [HH]
  int N, NC, C, H, W, HH, WW, align;
};

__device__ __forceinline__ float crop_coord(float sw, float ew, float start, float extent, int size, int align) {
  const float p0 = __fsub_rn(__fmul_rn(2.f, start), 1.f);
  const float p1 = __fsub_rn(__fmul_rn(2.f, __fadd_rn(start, extent)), 1.f);
  const float g = __fadd_rn(__fmul_rn(sw, p0), __fmul_rn(ew, p1));
  if (align) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
}

// backend 'jj' (bilinear.py:97-152, selected with align = 2): coordinates stay in [0, 1] (no 2x - 1), are scaled by the
// image size, and the four taps are floor / floor + 1 CLAMPED into the image with the weights (x1 - X), (X - x0)
// taken against the clamped taps (so a sample beyond the last pixel gets a zero or negative weight, as there)
struct JJAxis { int a, b; float wa, wb; };
__device__ __forceinline__ JJAxis jj_axis(float sw, float ew, float start, float extent, int size) {
  const float X = __fmul_rn(__fadd_rn(__fmul_rn(sw, start), __fmul_rn(ew, __fadd_rn(start, extent))), (float)size);
  const float hi = (float)(size - 1);
  const float a = fminf(fmaxf(floorf(X), 0.f), hi);
  const float b = fminf(fmaxf(__fadd_rn(a, 1.f), 0.f), hi);
  JJAxis r;
  r.a = (int)a; r.b = (int)b;
  r.wa = __fsub_rn(b, X); r.wb = __fsub_rn(X, a);
  return r;
}

__global__ void crop_img_kernel(const int* __restrict__ crop_off, int N, int* __restrict__ crop_img) {
  const int n = blockIdx.x;
  for (int i = crop_off[n] + threadIdx.x; i < crop_off[n + 1]; i += blockDim.x) crop_img[i] = n;
}

// crops[i, c, yy, xx]: one block per crop, threads over output pixels, channels innermost per thread
__global__ void __launch_bounds__(256) crop_fwd_kernel(CropParams p, const int* __restrict__ crop_img,
                                                       float* __restrict__ out) {
  const int i = blockIdx.x, n = crop_img[i];
  const float4 b = *reinterpret_cast<const float4*>(p.bbox + 4 * (size_t)i);
  const size_t plane = (size_t)p.H * p.W;
  const float* img = p.feats + (size_t)n * p.C * plane;
  for (int px = threadIdx.x + blockIdx.y * blockDim.x; px < p.HH * p.WW; px += blockDim.x * gridDim.y) {
    const int yy = px / p.WW, xx = px % p.WW;
    if (p.align == 2) {
      const JJAxis ax = jj_axis(p.swx[xx], p.ewx[xx], b.x, b.z, p.W), ay = jj_axis(p.swy[yy], p.ewy[yy], b.y, b.w, p.H);
      const float w1 = __fmul_rn(ax.wa, ay.wa), w2 = __fmul_rn(ax.wa, ay.wb), w3 = __fmul_rn(ax.wb, ay.wa), w4 = __fmul_rn(ax.wb, ay.wb);
      for (int c = 0; c < p.C; ++c) {
        const float* src = img + (size_t)c * plane;
        const float v1 = __ldg(src + (size_t)ay.a * p.W + ax.a), v2 = __ldg(src + (size_t)ay.b * p.W + ax.a);
        const float v3 = __ldg(src + (size_t)ay.a * p.W + ax.b), v4 = __ldg(src + (size_t)ay.b * p.W + ax.b);
        out[(((size_t)i * p.C + c) * p.HH + yy) * p.WW + xx] =
            __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
      }
      continue;
    }
    const float ix = crop_coord(p.swx[xx], p.ewx[xx], b.x, b.z, p.W, p.align);
    const float iy = crop_coord(p.swy[yy], p.ewy[yy], b.y, b.w, p.H, p.align);
    const float fx = floorf(ix), fy = floorf(iy);
    const float tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy);
    const bool okx = fx >= -1.f && fx < (float)p.W, oky = fy >= -1.f && fy < (float)p.H;
    const int x0 = okx ? (int)fx : -2, y0 = oky ? (int)fy : -2;
    const bool vx0 = x0 >= 0 && x0 < p.W, vx1 = x0 >= -1 && x0 < p.W - 1;
    const bool vy0 = y0 >= 0 && y0 < p.H, vy1 = y0 >= -1 && y0 < p.H - 1;
    const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
    for (int c = 0; c < p.C; ++c) {
      const float* src = img + (size_t)c * plane;
      const float vnw = (vy0 && vx0) ? __ldg(src + (size_t)y0 * p.W + x0) : 0.f;
      const float vne = (vy0 && vx1) ? __ldg(src + (size_t)y0 * p.W + x0 + 1) : 0.f;
      const float vsw = (vy1 && vx0) ? __ldg(src + (size_t)(y0 + 1) * p.W + x0) : 0.f;
      const float vse = (vy1 && vx1) ? __ldg(src + (size_t)(y0 + 1) * p.W + x0 + 1) : 0.f;
      out[(((size_t)i * p.C + c) * p.HH + yy) * p.WW + xx] = vnw * wnw + vne * wne + vsw * wsw + vse * wse;
    }
  }
}

// dfeats[n, c, y, x] = sum_{crops i of n} sum_{yy, xx} dcrops[i, c, yy, xx] * wy(yy -> y) * wx(xx -> x)
constexpr int BW_THREADS = 256;
constexpr int BW_CH = 4;

__global__ void __launch_bounds__(BW_THREADS) crop_bwd_kernel(CropParams p, const float* __restrict__ dcrops,
                                                              float* __restrict__ dfeats) {
  extern __shared__ float sm[];
  float* sx = sm;              // [WW] sample column of each crop column
  float* sy = sm + p.WW;       // [HH]
  const int n = blockIdx.y;
  const int pix = blockIdx.x * BW_THREADS + threadIdx.x;
  const bool live = pix < p.H * p.W;
  const int y = live ? pix / p.W : 0, x = live ? pix % p.W : 0;
  const int ymin = (blockIdx.x * BW_THREADS) / p.W, ymax = min(p.H * p.W - 1, blockIdx.x * BW_THREADS + BW_THREADS - 1) / p.W;
  const size_t plane = (size_t)p.H * p.W;
  for (int c0 = 0; c0 < p.C; c0 += BW_CH) {
    float acc[BW_CH];
#pragma unroll
    for (int j = 0; j < BW_CH; ++j) acc[j] = 0.f;
    for (int i = p.crop_off[n]; i < p.crop_off[n + 1]; ++i) {
      const float4 b = *reinterpret_cast<const float4*>(p.bbox + 4 * (size_t)i);
      __syncthreads();
      for (int k = threadIdx.x; k < p.WW + p.HH; k += BW_THREADS) {
        if (k < p.WW) sx[k] = crop_coord(p.swx[k], p.ewx[k], b.x, b.z, p.W, p.align);
        else sy[k - p.WW] = crop_coord(p.swy[k - p.WW], p.ewy[k - p.WW], b.y, b.w, p.H, p.align);
      }
      __syncthreads();
      if (!live) continue;
      if (p.align == 2) {
        // 'jj' taps: weight of image pixel (y, x) in crop pixel (yy, xx) = wy * wx with both clamped taps counted
        for (int yy = 0; yy < p.HH; ++yy) {
          const JJAxis ay = jj_axis(p.swy[yy], p.ewy[yy], b.y, b.w, p.H);
          const float wy = (ay.a == y ? ay.wa : 0.f) + (ay.b == y ? ay.wb : 0.f);
          if (ay.a != y && ay.b != y) continue;
          for (int xx = 0; xx < p.WW; ++xx) {
            const JJAxis ax = jj_axis(p.swx[xx], p.ewx[xx], b.x, b.z, p.W);
            if (ax.a != x && ax.b != x) continue;
            const float w = ((ax.a == x ? ax.wa : 0.f) + (ax.b == x ? ax.wb : 0.f)) * wy;
#pragma unroll
            for (int j = 0; j < BW_CH; ++j)
              if (c0 + j < p.C)
                acc[j] = fmaf(__ldg(dcrops + (((size_t)i * p.C + c0 + j) * p.HH + yy) * p.WW + xx), w, acc[j]);
          }
        }
        continue;
      }
      for (int yy = 0; yy < p.HH; ++yy) {
        const float iy = sy[yy];
        const float fy = floorf(iy);
        if (!(fy >= (float)(ymin - 1) && fy <= (float)ymax)) continue;       // block-uniform cull
        const float ty = __fsub_rn(iy, fy);
        const int y0 = (int)fy;
        const float wy = y0 == y ? 1.f - ty : (y0 + 1 == y ? ty : 0.f);
        if (!(y0 == y || y0 + 1 == y)) continue;
        for (int xx = 0; xx < p.WW; ++xx) {
          const float ix = sx[xx];
          const float fx = floorf(ix);
          if (!(fx >= (float)(x - 1) && fx <= (float)x)) continue;
          const float tx = __fsub_rn(ix, fx);
          const float wx = ((int)fx == x) ? 1.f - tx : tx;
          // ATen weights are products in this order: (x part) * (y part)
          const float w = wx * wy;
#pragma unroll
          for (int j = 0; j < BW_CH; ++j)
            if (c0 + j < p.C)
              acc[j] = fmaf(__ldg(dcrops + (((size_t)i * p.C + c0 + j) * p.HH + yy) * p.WW + xx), w, acc[j]);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < BW_CH; ++j)
        if (c0 + j < p.C) dfeats[((size_t)n * p.C + c0 + j) * plane + pix] = acc[j];
    }
  }
}

int fill(CropParams& p, const float* feats, const float* bbox, const int* crop_off, const float* swx, const float* ewx,
         const float* swy, const float* ewy, int N, int NC, int C, int H, int W, int HH, int WW, int align) {
  CSG_REQUIRE(N >= 0 && NC >= 0 && C > 0 && H > 0 && W > 0 && HH > 0 && WW > 0,
              "crop_bbox: bad sizes N=%d NC=%d C=%d H=%d W=%d HH=%d WW=%d", N, NC, C, H, W, HH, WW);
  CSG_REQUIRE((reinterpret_cast<uintptr_t>(bbox) & 15) == 0, "crop_bbox: bbox must be 16-byte aligned");
  p.feats = feats; p.bbox = bbox; p.crop_off = crop_off; p.swx = swx; p.ewx = ewx; p.swy = swy; p.ewy = ewy;
  p.N = N; p.NC = NC; p.C = C; p.H = H; p.W = W; p.HH = HH; p.WW = WW; p.align = align;
  return 0;
}

}  // namespace

CSG_API size_t csg_crop_bbox_workspace(int NC) { return (size_t)(NC + 1) * 4 + 64; }

CSG_API int csg_crop_bbox_fwd(const float* feats, const float* bbox, const int* crop_off, const float* swx,
                              const float* ewx, const float* swy, const float* ewy, float* crops, int N, int NC,
                              int C, int H, int W, int HH, int WW, int align_corners, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
  CropParams p;
  if (int rc = fill(p, feats, bbox, crop_off, swx, ewx, swy, ewy, N, NC, C, H, W, HH, WW, align_corners)) return rc;
  if (N == 0 || NC == 0) return 0;
  CSG_REQUIRE(workspace_bytes >= csg_crop_bbox_workspace(NC), "crop_bbox: workspace too small");
  int* crop_img = reinterpret_cast<int*>(workspace);
  crop_img_kernel<<<N, 64, 0, stream>>>(crop_off, N, crop_img);
  CSG_CHECK_LAUNCH("csg_crop_bbox_fwd crop_img");
  dim3 grid(NC, csg_div_up((long long)HH * WW, 1024));
  crop_fwd_kernel<<<grid, 256, 0, stream>>>(p, crop_img, crops);
  CSG_CHECK_LAUNCH("csg_crop_bbox_fwd");
  return 0;
}

CSG_API int csg_crop_bbox_bwd(const float* dcrops, const float* bbox, const int* crop_off, const float* swx,
                              const float* ewx, const float* swy, const float* ewy, float* dfeats, int N, int NC,
                              int C, int H, int W, int HH, int WW, int align_corners, cudaStream_t stream) {
  CropParams p;
  if (int rc = fill(p, nullptr, bbox, crop_off, swx, ewx, swy, ewy, N, NC, C, H, W, HH, WW, align_corners)) return rc;
  if (N == 0) return 0;
  dim3 grid(csg_div_up((long long)H * W, BW_THREADS), N);
  crop_bwd_kernel<<<grid, BW_THREADS, (size_t)(HH + WW) * 4, stream>>>(p, dcrops, dfeats);
  CSG_CHECK_LAUNCH("csg_crop_bbox_bwd");
  return 0;
}
