// bf16 tensor-core GEMM family for sm_100a: tcgen05.mma (cta_group::1, M=128, N<=256) with fp32
// accumulators in TMEM, operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte
// swizzle) or -- for the fused triple gather -- by cp.async row gathers written in the same swizzle.
//
// This is the throughput engine of net1 / net2 (sg2im/graph.py:33-41,67,110): the reference runs
// them as fp32 nn.Linear calls on a materialised [B, T, 3D] concat; here
//   F1   hidden = relu([obj[s] | pred | obj[o]] W1^T + b1)     A rows gathered straight into smem
//   F2   out    = relu(hidden W2^T + b2) * conf
//   dX-type  dy W   (weights pre-transposed once per step so both operands stay K-major)
//   dW-type  dy^T x (both operands MN-major, K = triples, split-K over persistent CTAs)
//
// Kernel anatomy (persistent, warp specialised, 320 threads):
//   warp 0      TMA producer (one lane)            smem ring of STAGES x (A 16 KB + B <= 32 KB)
//   warp 1      MMA issuer (one lane)              tcgen05.mma + tcgen05.commit -> empty / tmem_full
//   warps 2-5   epilogue: tcgen05.ld 32x32b.x32, bias / ReLU / row scale / ReLU mask, 16-byte stores
//   warps 6-9   gather producers (only in the gather variants)
// TMEM: 512 columns = 2 accumulator stages x 256, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 320;
constexpr int GATHER_THREADS = 128;
constexpr int ACC_STRIDE = 256;                   // TMEM columns per accumulator stage

enum { G_NONE = 0, G_A = 1, G_B = 2 };

struct TcParams {
  int M, N, K;                 // K = reduction length (rows of the MN-major operands)
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  void* C;                     // [splits][M][ldc] (splits > 1: fp32 partials)
  int ldc, out_f32;
  const float* bias;           // [N]
  int relu;
  const float* rowscale;       // [M]
  const __nv_bfloat16* mask_aux;   // [M, ld_aux]  multiply by (aux > 0)
  int ld_aux;
  // fused gather of [obj[s] | pred | obj[o]] rows
  const __nv_bfloat16* g_obj;
  const __nv_bfloat16* g_pred;
  const int* g_sidx;
  const int* g_oidx;
  int g_din, g_dp, g_ldp, g_rows;
};

// ------------------------------------------------------------------------------------------ PTX
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
// Bounded wait: a protocol bug traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("csg gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ const __nv_bfloat16* gather_src(const TcParams& p, int t, int c) {
  // element c of the virtual row [obj[s_t] | pred[t] | obj[o_t]]
  if (c < p.g_din) return p.g_obj + (size_t)p.g_sidx[t] * p.g_din + c;
  if (c < p.g_din + p.g_dp) return p.g_pred + (size_t)t * p.g_ldp + (c - p.g_din);
  return p.g_obj + (size_t)p.g_oidx[t] * p.g_din + (c - p.g_din - p.g_dp);
}

struct __align__(8) Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int BN>
struct Tile {
  static constexpr int B_STAGE = BN * BLOCK_K * 2;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * (A_STAGE_BYTES + B_STAGE) + sizeof(Barriers);
};

template <int BN, bool MN, int GATHER>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + STAGES * A_STAGE_BYTES;
  constexpr int B_STAGE = Tile<BN>::B_STAGE;
  Barriers* bars = reinterpret_cast<Barriers*>(base_ptr + (size_t)STAGES * (A_STAGE_BYTES + B_STAGE));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1 + (GATHER != G_NONE ? GATHER_THREADS : 0));
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    if (GATHER != G_A) tma_prefetch_desc(&tmA);
    if (GATHER != G_B) tma_prefetch_desc(&tmB);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  auto decode = [&](int tile, int& mt, int& nt, int& sp) {
    nt = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    mt = r % p.m_tiles;
    sp = r / p.m_tiles;
  };
  auto kb_range = [&](int sp, int& kb0, int& kb1) {
    kb0 = sp * p.kb_per_split;
    kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      int stage = 0, phase = 0;
      const bool loadA = GATHER != G_A, loadB = GATHER != G_B;
      const uint32_t bytes = (loadA ? A_STAGE_BYTES : 0) + (loadB ? B_STAGE : 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt, sp, kb0, kb1;
        decode(tile, mt, nt, sp);
        kb_range(sp, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&bars->full[stage]);
          mbar_arrive_expect_tx(fb, bytes);
          const uint32_t a_dst = sA + stage * A_STAGE_BYTES, b_dst = sB + stage * B_STAGE;
          if (!MN) {
            // K-major: rows = M (or N), 64 contiguous k elements per 128-byte row
            if (loadA) tma_load_2d(a_dst, &tmA, fb, kb * BLOCK_K, mt * BLOCK_M);
            if (loadB) tma_load_2d(b_dst, &tmB, fb, kb * BLOCK_K, nt * BN);
          } else {
            // MN-major: rows = k (64 per block), one [64 k x 64 mn] box per 64-wide chunk
            if (loadA)
              for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(a_dst + c * 8192, &tmA, fb, mt * BLOCK_M + c * 64, kb * BLOCK_K);
            if (loadB)
              for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, &tmB, fb, nt * BN + c * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BN, MN, MN);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        int mt, nt, sp, kb0, kb1;
        decode(tile, mt, nt, sp);
        kb_range(sp, kb0, kb1);
        const int as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bars->tmem_empty[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint32_t a_base = sA + stage * A_STAGE_BYTES, b_base = sB + stage * B_STAGE;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            uint64_t ad, bd;
            if (!MN) {
              ad = make_desc(a_base + k * 32, 16, 1024);
              bd = make_desc(b_base + k * 32, 16, 1024);
            } else {
              ad = make_desc(a_base + k * 2048, 8192, 1024);
              bd = make_desc(b_base + k * 2048, 8192, 1024);
            }
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&bars->empty[stage]));        // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(smem_u32(&bars->tmem_full[as]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0) umma_commit(smem_u32(&bars->tmem_full[as]));   // empty k range (never scheduled)
      }
    }
  } else if (warp < 6) {
    // ================================================================== epilogue
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int mt, nt, sp;
      decode(tile, mt, nt, sp);
      const int as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(smem_u32(&bars->tmem_full[as]), aphase);
      tc_fence_after();
      const int row = mt * BLOCK_M + q * 32 + lane;
      const bool rowok = row < p.M;
      const float rs = (p.rowscale && rowok) ? p.rowscale[row] : 1.f;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        const int col = nt * BN + c * 32;
        if (rowok && col < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.splits == 1) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.rowscale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= rs;
            }
            if (p.mask_aux) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.mask_aux + (size_t)row * p.ld_aux + col);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                uint4 a = __ldg(ap + j4);
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
                  if (!(lo != 0 && lo < 0x8000u)) v[j4 * 8 + e * 2] = 0.f;
                  if (!(hi != 0 && hi < 0x8000u)) v[j4 * 8 + e * 2 + 1] = 0.f;
                }
              }
            }
          }
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.C) + ((size_t)sp * p.M + row) * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4) st_f4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              uint4 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
              u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst + j) = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[as]));
    }
  } else if (GATHER != G_NONE) {
    // ================================================================== gather producers (warps 6-9)
    // Every thread copies one 128-byte row piece per stage with 8 x 16-byte cp.async, applying the
    // SWIZZLE_128B pattern TMA would have used: chunk c of row r lands at chunk (c ^ (r & 7)).
    // Completion is published LAG stages later: wait_group -> fence.proxy.async -> mbarrier.arrive.
    constexpr int LAG = 2;
    const int g = threadIdx.x - 6 * 32;            // 0..127
    int stage = 0, phase = 0;
    int pend_stage[LAG];
    int npend = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mt, nt, sp, kb0, kb1;
      decode(tile, mt, nt, sp);
      kb_range(sp, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        if (GATHER == G_A) {
          // A tile [128 triples x 64 k]: row g = triple mt*128+g, columns kb*64 .. +63 of the virtual input
          int t = min(mt * BLOCK_M + g, p.g_rows - 1);
          const __nv_bfloat16* src = gather_src(p, t, kb * BLOCK_K);
          const uint32_t dst = sA + stage * A_STAGE_BYTES + g * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async_16(dst + ((c ^ (g & 7)) << 4), src + c * 8);
        } else {
          // B tile, MN-major: BN/64 chunks of [64 triples x 64 n]; rows = triples kb*64 + r
          constexpr int ROWS = (BN / 64) * 64;      // row pieces per stage
          for (int i = g; i < ROWS; i += GATHER_THREADS) {
            int chunk = i >> 6, r = i & 63;
            int t = min(kb * BLOCK_K + r, p.g_rows - 1);
            const __nv_bfloat16* src = gather_src(p, t, nt * BN + chunk * 64);
            const uint32_t dst = sB + stage * B_STAGE + chunk * 8192 + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) cp_async_16(dst + ((c ^ (r & 7)) << 4), src + c * 8);
          }
        }
        cp_async_commit();
        if (npend == LAG) {
          cp_async_wait<LAG>();                    // the oldest outstanding group has landed
          fence_proxy_async();
          mbar_arrive(smem_u32(&bars->full[pend_stage[0]]));
#pragma unroll
          for (int i = 0; i + 1 < LAG; ++i) pend_stage[i] = pend_stage[i + 1];
          --npend;
        }
        pend_stage[npend++] = stage;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int i = 0; i < npend; ++i) mbar_arrive(smem_u32(&bars->full[pend_stage[i]]));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

__global__ void splitk_reduce_tc_kernel(const float* __restrict__ partial, float* __restrict__ C, long long MN, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[(size_t)z * MN + i];
  C[i] = acc;
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, `outer` rows of `row_bytes` pitch, box [64 x box_outer]
int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_bytes, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  CSG_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  CSG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (row_bytes & 15) == 0,
              "gemm_tc: operand pointer / pitch must be 16-byte aligned");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSG_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int BN, bool MN, int GATHER>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t stream) {
  constexpr size_t smem = Tile<BN>::SMEM;
  static bool configured = false;
  if (!configured) {
    CSG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = tiles < csg_num_sms() ? tiles : csg_num_sms();
  gemm_tc_kernel<BN, MN, GATHER><<<grid, NUM_THREADS, smem, stream>>>(ma, mb, p);
  CSG_CHECK_LAUNCH("csg_gemm_bf16");
  return 0;
}

template <bool MN, int GATHER>
int launch_bn(int BN, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t stream) {
  switch (BN) {
    case 64: return launch<64, MN, GATHER>(ma, mb, p, stream);
    case 128: return launch<128, MN, GATHER>(ma, mb, p, stream);
    case 192: return launch<192, MN, GATHER>(ma, mb, p, stream);
    case 256: return launch<256, MN, GATHER>(ma, mb, p, stream);
  }
  csg_set_error("gemm_tc: unsupported BLOCK_N %d", BN);
  return 1;
}

int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N >= 256) return 256;
  if (N > 128) return 192 >= N ? 192 : 256;
  if (N > 64) return 128;
  return 64;
}

}  // namespace

CSG_API size_t csg_gemm_bf16_workspace(int M, int N, int K, int mn_major) {
  if (!mn_major) return 0;
  return (size_t)64 * M * N * sizeof(float);
}

// mn_major = 0:  C[M,N] = epi(A[M,K] * B[N,K]^T)        A, B row-major with K contiguous (bf16)
//                gather = 1: A rows are [obj[s] | pred | obj[o]] (K = 2*Din + Dp), A / lda ignored
// mn_major = 1:  C[M,N] = A[K,M]^T * B[K,N]             A, B row-major with M / N contiguous; fp32 output,
//                K split over the persistent CTAs (workspace: csg_gemm_bf16_workspace bytes)
//                gather = 2: B rows are the gathered triple input (N = 2*Din + Dp), B / ldb ignored
// out_f32 selects fp32 or bf16 C.  bias [N] fp32, rowscale [M] fp32, mask_aux [M, ld_aux] bf16 may be null.
CSG_API int csg_gemm_bf16(int mn_major, int gather, int M, int N, int K,
                          const void* A, int lda, const void* B, int ldb, void* C, int ldc, int out_f32,
                          const float* bias, int relu, const float* rowscale, const void* mask_aux, int ld_aux,
                          const void* g_obj, const void* g_pred, const int* g_sidx, const int* g_oidx,
                          int g_din, int g_dp, int g_ldp,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (M == 0 || N == 0) return 0;
  CSG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad sizes M=%d N=%d K=%d", M, N, K);
  CSG_REQUIRE(N % 32 == 0, "gemm_bf16: N=%d must be a multiple of 32", N);
  CSG_REQUIRE((ldc % 8) == 0 || out_f32, "gemm_bf16: bf16 ldc must be a multiple of 8");
  CSG_REQUIRE(!out_f32 || (ldc % 4) == 0, "gemm_bf16: fp32 ldc must be a multiple of 4");
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.out_f32 = out_f32;
  p.bias = bias; p.relu = relu; p.rowscale = rowscale;
  p.mask_aux = reinterpret_cast<const __nv_bfloat16*>(mask_aux); p.ld_aux = ld_aux;
  p.g_obj = reinterpret_cast<const __nv_bfloat16*>(g_obj); p.g_pred = reinterpret_cast<const __nv_bfloat16*>(g_pred);
  p.g_sidx = g_sidx; p.g_oidx = g_oidx; p.g_din = g_din; p.g_dp = g_dp; p.g_ldp = g_ldp;
  p.g_rows = mn_major ? K : M;
  if (gather) {
    CSG_REQUIRE(g_obj && g_pred && g_sidx && g_oidx, "gemm_bf16: gather sources missing");
    CSG_REQUIRE(g_din % 64 == 0 && g_dp % 64 == 0 && g_ldp % 8 == 0, "gemm_bf16: gather dims must be multiples of 64");
    CSG_REQUIRE((gather == 1 && !mn_major && K == 2 * g_din + g_dp) || (gather == 2 && mn_major && N == 2 * g_din + g_dp),
                "gemm_bf16: gather mode / shape mismatch");
  }
  if (p.mask_aux) CSG_REQUIRE(ld_aux % 8 == 0, "gemm_bf16: ld_aux must be a multiple of 8");
  int BN = pick_bn(N);
  if (gather == 2) {   // a gathered 64-column chunk must not straddle two source segments: guaranteed by %64 dims
    BN = (N % 192 == 0) ? 192 : (N % 128 == 0 ? 128 : 64);
  }
  p.m_tiles = csg_div_up(M, BLOCK_M);
  p.n_tiles = csg_div_up(N, BN);
  p.kb_total = csg_div_up(K, BLOCK_K);
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  void* out = C;
  if (!mn_major) {
    CSG_REQUIRE(K % 8 == 0, "gemm_bf16: K=%d must be a multiple of 8", K);
    if (gather != 1) { if (int rc = make_map(&ma, A, K, M, (uint64_t)lda * 2, BLOCK_M)) return rc; }
    if (int rc = make_map(&mb, B, K, N, (uint64_t)ldb * 2, BN)) return rc;
  } else {
    CSG_REQUIRE(out_f32, "gemm_bf16: MN-major (weight-gradient) GEMMs write fp32");
    CSG_REQUIRE(!bias && !relu && !rowscale && !mask_aux, "gemm_bf16: MN-major GEMMs have no epilogue");
    int tiles = p.m_tiles * p.n_tiles;
    int splits = csg_div_up(csg_num_sms(), tiles);
    if (splits > 64) splits = 64;
    if (splits > p.kb_total) splits = p.kb_total;
    p.kb_per_split = csg_div_up(p.kb_total, splits);
    p.splits = csg_div_up(p.kb_total, p.kb_per_split);
    if (p.splits > 1) {
      CSG_REQUIRE(ldc == N, "gemm_bf16: split-K output must be contiguous");
      CSG_REQUIRE(workspace && workspace_bytes >= (size_t)p.splits * M * N * sizeof(float), "gemm_bf16: workspace too small");
      p.C = workspace;
    }
    if (int rc = make_map(&ma, A, M, K, (uint64_t)lda * 2, 64)) return rc;
    if (gather != 2) { if (int rc = make_map(&mb, B, N, K, (uint64_t)ldb * 2, 64)) return rc; }
  }
  int rc;
  if (!mn_major) rc = gather == 1 ? launch_bn<false, G_A>(BN, ma, mb, p, stream) : launch_bn<false, G_NONE>(BN, ma, mb, p, stream);
  else rc = gather == 2 ? launch_bn<true, G_B>(BN, ma, mb, p, stream) : launch_bn<true, G_NONE>(BN, ma, mb, p, stream);
  if (rc) return rc;
  if (p.splits > 1) {
    long long MN = (long long)M * N;
    splitk_reduce_tc_kernel<<<csg_div_up(MN, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(p.C),
                                                                     reinterpret_cast<float*>(out), MN, p.splits);
    CSG_CHECK_LAUNCH("csg_gemm_bf16 split-K reduce");
  }
  return 0;
}
