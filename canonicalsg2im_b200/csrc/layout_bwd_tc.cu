// Tensor-core backward of the layout compositor wrt the object vectors (sm_100a, tcgen05 kind::tf32).
//
//   dvecs[o, d] = sum_px S_o(px) * dout[n, d, px]          (autograd of sg2im/layout.py:12-77 wrt `vecs`)
//
// is, per image, the GEMM  D[128 channels x N objects] = A[128 x HW] * B[N x HW]^T  whose A operand is the incoming
// gradient exactly as it lies in HBM (pixels contiguous = K-major), so it is streamed once by TMA (128-byte swizzle,
// 32 pixels x 128 channels = 16 KB per tile) and never touched by a per-object loop; B (the pixel weights of the
// image's objects for those 32 pixels) is generated on the fly in shared memory from the per-object factor tables.
// The accumulator lives in TMEM for a whole image.
//
// fp32 contract (1e-5): a tf32 operand keeps the top 19 bits of its 32-bit container (the low 13 are ignored), so both
// operands are split x = hi + lo (lo = the exact fp32 remainder) and three products are accumulated per 8-pixel step,
// hi*hi + hi*lo + lo*hi (the dropped lo*lo term is < 2^-22 relative).  To spare shared-memory bandwidth -- the
// binding resource of this kernel -- the raw gradient tile serves as its own hi part (never rewritten), and B is
// staged as one [hi rows ; lo rows] operand: one MMA of N = 2*NB gives A_hi*B_hi and A_hi*B_lo in separate TMEM
// columns, a second of N = NB adds A_lo*B_hi; the epilogue adds the two column groups.
//
// Work split: the N_img * (D / 128) * (HW / 32) tiles are dealt out in equal contiguous runs to one persistent CTA
// per SM (the contraction is dense, so equal bytes = equal time); a run may span image boundaries, at which the
// accumulator is flushed to a per-(image, CTA-rank) partial slot; a second kernel adds the slots in CTA order.
//
// Warps: 0 = TMA producer, 1 = MMA issuer, 2-9 = operand preparation (generation of B while the tile's TMA load is
// in flight, then the split of A); warps 2-5 also flush the accumulator at image boundaries.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

namespace {

constexpr int TC_M = 128;            // channels per accumulator (TMEM lanes)
constexpr int TC_KT = 32;            // pixels per tile: 32 fp32 = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_M * TC_KT * 4;      // 16 KB
constexpr int TC_MAX_STAGES = 10;    // raw gradient ring (16 KB per stage): deep, it has to cover the DRAM latency
constexpr int TC_LO_STAGES = 3;      // [A lo | B hi ; B lo] ring: short-lived, between the preparation warps and the MMAs
constexpr int TC_PREP = 256;         // preparation threads (warps 2-9; warps 2-5 also run the epilogue)
constexpr int TC_THREADS = 64 + TC_PREP;
constexpr int TC_NMAX = 64;          // objects per image handled by this path

struct TcBwdParams {
  const float* axg;       // [NO][W]  boxes: column factor; masks: column sample coordinate
  const float* ayg;       // [NO][H]
  const float* masks;     // [NO][S][S] or nullptr
  const int* obj_off;     // [N + 1]
  float* partial;         // [max_slots][NO][D]
  int N, NO, D, H, W, S;
  int cblocks;            // D / 128
  int tpi;                // tiles per (image, channel block) = H * W / 32
  int total_tiles, tpc;   // all tiles; tiles per CTA
  int NB;                 // MMA N = rows of the B tile (multiple of 16, >= max objects per image)
  int stages;
  int debug;              // scratch/bench_layout.py: 1 = no operand preparation, 2 = no MMAs (wrong results)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void tc_timeout(uint32_t bar, uint32_t parity) {
  printf("csg layout_bwd_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
// Bounded wait: a protocol bug traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) tc_timeout(bar, parity);
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory descriptor (cute::UMMA::SmemDescriptor), as in gemm_tc.cu
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor: D = f32 [4,6) = 1, A = B = tf32 [7,10) = [10,13) = 2, both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// coordinate -> bilinear taps, and the 4-tap read of an S x S mask with zero padding (layout_common.cuh's
// coord_tap / mask_weight; repeated here because this translation unit does not need the rest of that header)
struct TapT { int i0; float w0, w1; };
__device__ __forceinline__ TapT tap_of(float ix, int size) {
  const float f = floorf(ix);
  const float t = __fsub_rn(ix, f);
  TapT r;
  r.w0 = __fsub_rn(1.f, t);
  r.w1 = t;
  r.i0 = (f >= -2.f && f <= (float)size) ? (int)f : -2;
  return r;
}
__device__ __forceinline__ float mask_read(const float* __restrict__ m, int S, const TapT& tx, const TapT& ty) {
  const int ix = tx.i0, iy = ty.i0;
  const bool vx0 = ix >= 0 && ix < S, vx1 = ix >= -1 && ix < S - 1;
  const bool vy0 = iy >= 0 && iy < S, vy1 = iy >= -1 && iy < S - 1;
  const float m00 = (vy0 && vx0) ? __ldg(m + iy * S + ix) : 0.f;
  const float m01 = (vy0 && vx1) ? __ldg(m + iy * S + ix + 1) : 0.f;
  const float m10 = (vy1 && vx0) ? __ldg(m + (iy + 1) * S + ix) : 0.f;
  const float m11 = (vy1 && vx1) ? __ldg(m + (iy + 1) * S + ix + 1) : 0.f;
  return m00 * (tx.w0 * ty.w0) + m01 * (tx.w1 * ty.w0) + m10 * (tx.w0 * ty.w1) + m11 * (tx.w1 * ty.w1);
}

struct __align__(8) TcBars {
  uint64_t full[TC_MAX_STAGES];     // TMA -> preparation warps
  uint64_t empty[TC_MAX_STAGES];    // MMA retired -> TMA producer (raw ring)
  uint64_t ready[TC_LO_STAGES];     // preparation warps -> MMA issuer
  uint64_t lo_empty[TC_LO_STAGES];  // MMA retired -> preparation warps (lo / B ring)
  uint64_t acc_full, acc_empty;     // MMA issuer <-> epilogue
  uint32_t tmem_base;
};

// shared-memory plan: [A raw (= hi): stages x 16 KB][A lo: 3 x 16 KB][B (hi rows ; lo rows): 3 x 2*NB*128][barriers]
__host__ __device__ inline size_t tc_smem_bytes(int stages, int NB) {
  return (size_t)stages * TC_A_BYTES + (size_t)TC_LO_STAGES * (TC_A_BYTES + 2 * NB * 128) + sizeof(TcBars);
}

template <bool HAS_MASK>
__global__ void __launch_bounds__(TC_THREADS, 1)
layout_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmG, const TcBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const int S = p.stages;
  const uint32_t sAhi = base, sAlo = base + S * TC_A_BYTES;
  const uint32_t b_bytes = (uint32_t)p.NB * 256;                    // hi rows then lo rows of one stage
  const uint32_t sB = sAlo + TC_LO_STAGES * TC_A_BYTES;
  TcBars* bars = reinterpret_cast<TcBars*>(smem_raw + (size_t)S * TC_A_BYTES + (size_t)TC_LO_STAGES * (TC_A_BYTES + b_bytes));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("csg layout_bwd_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < TC_LO_STAGES; ++s) {
      mbar_init(smem_u32(&bars->ready[s]), TC_PREP);
      mbar_init(smem_u32(&bars->lo_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_empty), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  const int t0 = blockIdx.x * p.tpc, t1 = min(p.total_tiles, t0 + p.tpc);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: one 16 KB gradient tile per stage
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int grp = t / p.tpi, kt = t % p.tpi;
        const int n = grp / p.cblocks, cblk = grp % p.cblocks;
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&bars->full[stage]);
        mbar_arrive_expect_tx(fb, TC_A_BYTES);
        tma_load_2d(sAhi + stage * TC_A_BYTES, &tmG, fb, kt * TC_KT, n * p.D + cblk * TC_M);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc2 = make_idesc_tf32(2 * p.NB), idesc1 = make_idesc_tf32(p.NB);
      const int sets = p.NB <= 32 ? 4 : 2;
      int stage = 0, phase = 0, seg = 0, ls = 0, lphase = 0;
      for (int t = t0; t < t1; ++t) {
        const int kt = t % p.tpi;
        const bool first = (t == t0) || kt == 0;          // first tile of an (image, channel block) segment in this CTA
        if (first) {
          mbar_wait(smem_u32(&bars->acc_empty), (seg & 1) ^ 1);    // the previous segment's sums have left TMEM
          tc_fence_after();
        }
        mbar_wait(smem_u32(&bars->ready[ls]), lphase);
        tc_fence_after();
        const uint32_t ahi = sAhi + stage * TC_A_BYTES, alo = sAlo + ls * TC_A_BYTES;
        const uint32_t bb = sB + ls * b_bytes;
        // Back-to-back MMAs into the same TMEM columns serialise on the accumulator (a narrow MMA is much shorter than
        // the accumulate latency), so the 8-pixel steps of a tile rotate over `sets` independent accumulator sets of
        // 3*NB columns [hi*hi | hi*lo | lo*hi]; the epilogue adds them up.
#pragma unroll
        for (int k = 0; k < TC_KT / 8; ++k) {
          const uint64_t dah = make_desc(ahi + k * 32), dal = make_desc(alo + k * 32), db = make_desc(bb + k * 32);
          if (p.debug & 2) continue;
          const uint32_t dset = tmem + (uint32_t)((k % sets) * 3 * p.NB);
          const uint32_t acc = (first && k < sets) ? 0u : 1u;
          umma_tf32(dset, dah, db, idesc2, acc);                    // [0, NB) += hi*hi, [NB, 2NB) += hi*lo
          umma_tf32(dset + 2 * p.NB, dal, db, idesc1, acc);         // [2NB, 3NB) += lo*hi
        }
        umma_commit(smem_u32(&bars->empty[stage]));
        umma_commit(smem_u32(&bars->lo_empty[ls]));
        const bool last = (t == t1 - 1) || kt == p.tpi - 1;
        if (last) { umma_commit(smem_u32(&bars->acc_full)); ++seg; }
        if (++stage == S) { stage = 0; phase ^= 1; }
        if (++ls == TC_LO_STAGES) { ls = 0; lphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ preparation + epilogue warps
    const int pt = threadIdx.x - 64;                    // 0..255
    const int q = warp & 3;                             // TMEM lane quarter of this warp
    int stage = 0, phase = 0, seg = 0, ls = 0, lphase = 0;
    for (int t = t0; t < t1; ++t) {
      const int grp = t / p.tpi, kt = t % p.tpi;
      const int n = grp / p.cblocks, cblk = grp % p.cblocks;
      const int obeg = p.obj_off[n], On = p.obj_off[n + 1] - obeg;
      const int pix0 = kt * TC_KT, y = pix0 / p.W, x0 = pix0 % p.W;
      // ---- B tile: weights of the image's objects for these 32 pixels, split hi / lo, written in the 128-byte swizzle
      // (16-byte piece j of row o sits at j ^ (o & 7)); rows >= On are zero.  Needs only the stage's buffers to be
      // free (the MMAs that last read them have retired), not the gradient tile, whose TMA load is still in flight.
      mbar_wait(smem_u32(&bars->lo_empty[ls]), lphase ^ 1);
      if (p.debug & 1) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
      } else {
        uint8_t* bh = smem_raw + (sB - base) + (size_t)ls * b_bytes;
        uint8_t* bl = bh + (size_t)p.NB * 128;           // NB is a multiple of 8: row NB + o swizzles like row o
        for (int i = pt; i < p.NB * 8; i += TC_PREP) {
          const int o = i >> 3, j = i & 7;              // object row, 16-byte piece (4 pixels)
          float w[4] = {0.f, 0.f, 0.f, 0.f};
          if (o < On) {
            const float* ax = p.axg + (size_t)(obeg + o) * p.W + x0 + 4 * j;
            const float ay = __ldg(p.ayg + (size_t)(obeg + o) * p.H + y);
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(ax));
            if (HAS_MASK) {
              const TapT ty = tap_of(ay, p.S);
              const float* m = p.masks + (size_t)(obeg + o) * p.S * p.S;
              w[0] = mask_read(m, p.S, tap_of(a4.x, p.S), ty);
              w[1] = mask_read(m, p.S, tap_of(a4.y, p.S), ty);
              w[2] = mask_read(m, p.S, tap_of(a4.z, p.S), ty);
              w[3] = mask_read(m, p.S, tap_of(a4.w, p.S), ty);
            } else {
              w[0] = a4.x * ay; w[1] = a4.y * ay; w[2] = a4.z * ay; w[3] = a4.w * ay;
            }
          }
          float4 h4, l4;
          split_tf32(w[0], h4.x, l4.x); split_tf32(w[1], h4.y, l4.y);
          split_tf32(w[2], h4.z, l4.z); split_tf32(w[3], h4.w, l4.w);
          const uint32_t off = (uint32_t)o * 128 + (uint32_t)((j ^ (o & 7)) << 4);
          *reinterpret_cast<float4*>(bh + off) = h4;
          *reinterpret_cast<float4*>(bl + off) = l4;
        }
        // ---- A tile: the exact remainder lo = x - hi(x) next to the raw tile (element-wise: the swizzle is irrelevant)
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        const float4* ah = reinterpret_cast<const float4*>(smem_raw + (size_t)stage * TC_A_BYTES);
        float4* al = reinterpret_cast<float4*>(smem_raw + (sAlo - base) + (size_t)ls * TC_A_BYTES);
#pragma unroll
        for (int u = 0; u < TC_A_BYTES / 16 / TC_PREP; ++u) {
          const int i = pt + u * TC_PREP;
          const float4 v = ah[i];
          float4 h4, l4;
          split_tf32(v.x, h4.x, l4.x); split_tf32(v.y, h4.y, l4.y);
          split_tf32(v.z, h4.z, l4.z); split_tf32(v.w, h4.w, l4.w);
          al[i] = l4;
        }
      }
      fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
      mbar_arrive(smem_u32(&bars->ready[ls]));
      if (++stage == S) { stage = 0; phase ^= 1; }
      if (++ls == TC_LO_STAGES) { ls = 0; lphase ^= 1; }

      // ---- end of a segment: flush the accumulator to this CTA's slot of the image
      const bool last = (t == t1 - 1) || kt == p.tpi - 1;
      if (last && warp < 6) {
        mbar_wait(smem_u32(&bars->acc_full), seg & 1);
        tc_fence_after();
        const int slot = blockIdx.x - (grp * p.tpi) / p.tpc;
        float* dst = p.partial + ((size_t)slot * p.NO + obeg) * p.D + cblk * TC_M + q * 32 + lane;
        const int sets = p.NB <= 32 ? 4 : 2;
        for (int c0 = 0; c0 < p.NB && c0 < On; c0 += 32) {
          float sum[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = 0.f;
          for (int g = 0; g < 3 * sets; ++g) {                       // fixed order: set 0 [hi*hi, hi*lo, lo*hi], set 1, ...
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.NB + c0), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < On && c0 + j < p.NB) dst[(size_t)(c0 + j) * p.D] = sum[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty));
      }
      if (last) ++seg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

__global__ void layout_bwd_tc_sum_kernel(const float* __restrict__ partial, const int* __restrict__ obj_off, int N,
                                         float* __restrict__ dvecs, int NO, int D, int cblocks, int tpi, int tpc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)NO * D) return;
  const int o = (int)(i / D), d = (int)(i % D);
  int lo = 0, hi = N;                      // image of object o: last n with obj_off[n] <= o
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (obj_off[mid] <= o) lo = mid; else hi = mid;
  }
  const long long g0 = (long long)(lo * cblocks + d / TC_M) * tpi;
  const int first = (int)(g0 / tpc), last = (int)((g0 + tpi - 1) / tpc);
  float acc = 0.f;
  for (int c = first; c <= last; ++c) acc += partial[((size_t)(c - first) * NO + o) * D + d];
  dvecs[i] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct TcPlan { int grid, tpc, max_slots, NB, stages; };
TcPlan tc_plan(int N, int D, int H, int W, int max_objs) {
  TcPlan r;
  const int tpi = (H * W) / TC_KT;
  const long long total = (long long)N * (D / TC_M) * tpi;
  long long grid = csg_num_sms();
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  r.tpc = (int)((total + grid - 1) / grid);
  r.grid = (int)((total + r.tpc - 1) / r.tpc);
  r.max_slots = (tpi + r.tpc - 2) / r.tpc + 1;
  if (r.max_slots < 1) r.max_slots = 1;
  int nb = ((max_objs > 0 ? max_objs : 16) + 15) & ~15;
  if (nb < 16) nb = 16;
  r.NB = nb;
  r.stages = TC_MAX_STAGES;
  while (r.stages > 2 && tc_smem_bytes(r.stages, nb) > 226 * 1024) --r.stages;
  return r;
}

}  // namespace

// Entry points used by csg_layout_bwd_vecs (layout.cu); the per-object tables come from layout_tables_kernel.
bool csg_layout_bwd_tc_eligible(const float* dout, int N, int D, int H, int W, int max_objs) {
  if (N <= 0 || max_objs <= 0 || max_objs > TC_NMAX) return false;
  if ((D % TC_M) != 0 || (W % TC_KT) != 0 || (reinterpret_cast<uintptr_t>(dout) & 15) != 0) return false;
  if ((long long)N * D > 0x7fffffffLL || (long long)H * W > 0x7fffffffLL) return false;
  return tc_get_encode() != nullptr;
}

size_t csg_layout_bwd_tc_partial_floats(int N, int NO, int D, int H, int W, int max_objs) {
  return (size_t)tc_plan(N, D, H, W, max_objs).max_slots * NO * D;
}

int csg_layout_bwd_tc_launch(const float* dout, const float* masks, const int* obj_off, const float* axg, const float* ayg,
                             float* partial, float* dvecs, int N, int NO, int D, int H, int W, int M, int max_objs,
                             cudaStream_t stream) {
  const TcPlan plan = tc_plan(N, D, H, W, max_objs);
  TcBwdParams p;
  p.axg = axg; p.ayg = ayg; p.masks = masks; p.obj_off = obj_off; p.partial = partial;
  p.N = N; p.NO = NO; p.D = D; p.H = H; p.W = W; p.S = masks ? M : 8;
  p.cblocks = D / TC_M;
  p.tpi = (H * W) / TC_KT;
  p.total_tiles = N * p.cblocks * p.tpi;
  p.tpc = plan.tpc;
  p.NB = plan.NB;
  p.stages = plan.stages;
  { const char* dbg = getenv("CSG_LBT_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  EncodeTiledFn enc = tc_get_encode();
  CSG_REQUIRE(enc != nullptr, "layout_bwd_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)H * W, (cuuint64_t)N * D};
  cuuint64_t strides[1] = {(cuuint64_t)H * W * 4};
  cuuint32_t box[2] = {TC_KT, TC_M};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dout), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSG_REQUIRE(r == CUDA_SUCCESS, "layout_bwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  const size_t smem = tc_smem_bytes(plan.stages, plan.NB);
  if (masks) {
    CSG_CUDA(cudaFuncSetAttribute(layout_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_bwd_tc_kernel<true><<<plan.grid, TC_THREADS, smem, stream>>>(map, p);
  } else {
    CSG_CUDA(cudaFuncSetAttribute(layout_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_bwd_tc_kernel<false><<<plan.grid, TC_THREADS, smem, stream>>>(map, p);
  }
  CSG_CHECK_LAUNCH("csg_layout_bwd_vecs tc");
  const long long n = (long long)NO * D;
  layout_bwd_tc_sum_kernel<<<csg_div_up(n, 256), 256, 0, stream>>>(partial, obj_off, N, dvecs, NO, D, p.cblocks, p.tpi, p.tpc);
  CSG_CHECK_LAUNCH("csg_layout_bwd_vecs tc sum");
  return 0;
}
