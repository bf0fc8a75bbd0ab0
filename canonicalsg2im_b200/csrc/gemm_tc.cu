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
// Kernel anatomy (persistent, warp specialised, 192 threads):
//   warp 0      producer: TMA tile loads (lane 0) and, in the gather variants, one TMA tile::gather4 per lane
//               (4 gathered rows of 128 bytes each, written in the same 128-byte swizzle) into the smem ring
//   warp 1      MMA issuer (one lane)              tcgen05.mma + tcgen05.commit -> empty / tmem_full
//   warps 2-5   epilogue: tcgen05.ld 32x32b.x32 -> bias / ReLU / row scale / ReLU mask -> bf16 -> swizzled smem
//               staging -> TMA store (each warp owns its 32 rows: no cross-warp barrier); the ReLU-mask operand
//               is prefetched by TMA into smem as well, so the epilogue issues no strided global accesses
// TMEM: 512 columns = 2 accumulator stages x 256, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int MAX_STAGES = 6;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;              // producer, MMA, 4 epilogue warps
constexpr int GATHER_WARPS = 4;               // extra warps 6-9 of the gather variants (8 issuing lanes each)
constexpr int ACC_STRIDE = 256;                   // TMEM columns per accumulator stage
constexpr int EPI_WARPS = 4;
constexpr int CHUNK_N = 64;                       // epilogue chunk: 64 bf16 columns = one 128-byte swizzle row
constexpr int CHUNK_BYTES = 32 * CHUNK_N * 2;     // per-warp staging buffer: 32 rows x 128 bytes
constexpr size_t SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA

enum { G_NONE = 0, G_A = 1, G_B = 2 };

struct TcParams {
  int M, N, K;                 // K = reduction length (rows of the MN-major operands)
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int stages, cbufs;           // smem ring depth; staging buffers per epilogue warp (1 or 2)
  int tma_epi;                 // bf16 output through swizzled smem + TMA store (K-major kernels)
  void* C;                     // [splits][M][ldc] (splits > 1: fp32 partials)
  int ldc, out_f32;
  const float* bias;           // [N]
  int relu;
  const float* rowscale;       // [M]
  const __nv_bfloat16* mask_aux;   // [M, ld_aux]  multiply by (aux > 0)
  int ld_aux;
  // fused gather of [obj[s] | pred | obj[o]] rows
  const int* g_sidx;
  const int* g_oidx;
  int g_din, g_dp, g_rows;
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
// tile::gather4: four rows (r0..r3) x 64 columns starting at column c0 of a 2-D tensor (box {64, 1}); the rows land
// at dst + {0, 128, 256, 384} bytes, swizzled by the destination address like a plain tile load.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
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

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t aux_full[EPI_WARPS][2];
  uint32_t tmem_base;
};

// Shared-memory plan (offsets from the 1024-byte aligned base):
//   [A ring: stages x 16 KB][B ring: stages x BN*128 B][C staging: 4 warps x cbufs x 4 KB]
//   [aux staging: 4 warps x 2 x 4 KB][bias: 4 warps x BN floats][barriers]
struct SmemPlan {
  uint32_t b, c, aux, bias, bars, total;
};
__host__ __device__ inline SmemPlan smem_plan(int BN, int stages, int cbufs, bool tma_epi, bool has_aux) {
  SmemPlan s;
  s.b = (uint32_t)stages * A_STAGE_BYTES;
  s.c = s.b + (uint32_t)stages * BN * 128;
  s.aux = s.c + (tma_epi ? EPI_WARPS * cbufs * CHUNK_BYTES : 0);
  s.bias = s.aux + ((tma_epi && has_aux) ? EPI_WARPS * 2 * CHUNK_BYTES : 0);
  s.bars = s.bias + (tma_epi ? EPI_WARPS * BN * 4 : 0);
  s.total = s.bars + (uint32_t)sizeof(Barriers) + 1024;      // + alignment slack
  return s;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// bf16 > 0  <=>  sign bit clear and magnitude non-zero
__device__ __forceinline__ bool bf16_pos(uint32_t h) { return h != 0 && h < 0x8000u; }

// tmA/tmB: operand maps (in the gather variants the gathered operand's slot holds the object-row table,
// box {64, 1}); tmP: predicate rows of the fused gather; tmC/tmX: output / ReLU-mask operand, box {64, 32}.
template <int BN, bool MN, int GATHER>
__global__ void __launch_bounds__(NUM_THREADS + (GATHER != G_NONE ? GATHER_WARPS * 32 : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmX, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  constexpr int B_STAGE = BN * BLOCK_K * 2;
  const bool has_aux = p.mask_aux != nullptr;
  const SmemPlan plan = smem_plan(BN, p.stages, p.cbufs, p.tma_epi != 0, has_aux);
  const uint32_t sA = base, sB = base + plan.b;
  Barriers* bars = reinterpret_cast<Barriers*>(base_ptr + plan.bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int STAGES = p.stages;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), EPI_WARPS);
    }
    for (int w = 0; w < EPI_WARPS; ++w)
      for (int s = 0; s < 2; ++s) mbar_init(smem_u32(&bars->aux_full[w][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (GATHER != G_NONE) tma_prefetch_desc(&tmP);
    if (p.tma_epi) tma_prefetch_desc(&tmC);
    if (p.tma_epi && has_aux) tma_prefetch_desc(&tmX);
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
    // ================================================================== producer (one lane): tile loads
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt, sp, kb0, kb1;
        decode(tile, mt, nt, sp);
        kb_range(sp, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&bars->full[stage]);
          // the whole stage is accounted here; bytes gathered by warps 6-9 may land before or after this arrive
          mbar_arrive_expect_tx(fb, A_STAGE_BYTES + B_STAGE);
          const uint32_t a_dst = sA + stage * A_STAGE_BYTES, b_dst = sB + stage * B_STAGE;
          if (!MN) {
            // K-major: rows = M (or N), 64 contiguous k elements per 128-byte row
            if (GATHER == G_NONE) tma_load_2d(a_dst, &tmA, fb, kb * BLOCK_K, mt * BLOCK_M);
            else {
              const int c0 = kb * BLOCK_K;     // predicate block of the virtual row [obj[s] | pred | obj[o]]
              if (c0 >= p.g_din && c0 < p.g_din + p.g_dp) tma_load_2d(a_dst, &tmP, fb, c0 - p.g_din, mt * BLOCK_M);
            }
            tma_load_2d(b_dst, &tmB, fb, kb * BLOCK_K, nt * BN);
          } else {
            // MN-major: rows = k (64 per block), one [64 k x 64 mn] box per 64-wide chunk
            for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(a_dst + c * 8192, &tmA, fb, mt * BLOCK_M + c * 64, kb * BLOCK_K);
            for (int c = 0; c < BN / 64; ++c) {
              const int n0 = nt * BN + c * 64;
              if (GATHER == G_NONE) tma_load_2d(b_dst + c * 8192, &tmB, fb, n0, kb * BLOCK_K);
              else if (n0 >= p.g_din && n0 < p.g_din + p.g_dp) tma_load_2d(b_dst + c * 8192, &tmP, fb, n0 - p.g_din, kb * BLOCK_K);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 6) {
    // ================================================================== gather producers (warps 6-9, lanes 0-7)
    // Object rows of the virtual operand arrive by TMA tile::gather4 (4 rows x 128 bytes per instruction, written
    // in the same 128-byte swizzle as a tile load).  A TMA instruction takes warp-uniform operands, so a warp
    // issues them one lane at a time: the work is spread over 4 warps x 8 lanes.
    if (GATHER != G_NONE && lane < 8) {
      const int gl = (warp - 6) * 8 + lane;          // 0..31
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt, sp, kb0, kb1;
        decode(tile, mt, nt, sp);
        kb_range(sp, kb0, kb1);
        if (GATHER == G_A) {
          // A tile [128 triples x 64 k]: this lane owns rows 4*gl .. 4*gl+3 (tmA = object table)
          int si[4], oi[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int t = mt * BLOCK_M + 4 * gl + i;
            const bool ok = t < p.g_rows;
            si[i] = ok ? __ldg(p.g_sidx + t) : 0;
            oi[i] = ok ? __ldg(p.g_oidx + t) : 0;
          }
          for (int kb = kb0; kb < kb1; ++kb) {
            const int c0 = kb * BLOCK_K;
            const bool seg_s = c0 < p.g_din, seg_o = c0 >= p.g_din + p.g_dp;
            if (seg_s || seg_o) {
              mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
              const uint32_t fb = smem_u32(&bars->full[stage]);
              const uint32_t dst = sA + stage * A_STAGE_BYTES + gl * 512;
              if (seg_s) tma_gather4(dst, &tmA, fb, c0, si[0], si[1], si[2], si[3]);
              else tma_gather4(dst, &tmA, fb, c0 - p.g_din - p.g_dp, oi[0], oi[1], oi[2], oi[3]);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          // B tile, MN-major: BN/64 chunks of [64 triples x 64 n], 16 row groups each (tmB = object table).
          // Job j = chunk * 16 + group; this lane owns jobs gl and gl + 32.
          constexpr int JOBS = (BN / 64) * 16;
          constexpr int PER = (JOBS + 31) / 32;
          int nx[PER][4];
          auto load_idx = [&](int kb, int (&dst)[PER][4]) {
#pragma unroll
            for (int u = 0; u < PER; ++u) {
              const int j = gl + 32 * u;
              const int n0 = nt * BN + (j >> 4) * 64;
              const bool seg_s = n0 < p.g_din, seg_o = n0 >= p.g_din + p.g_dp;
              const int* idx = seg_s ? p.g_sidx : p.g_oidx;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int t = kb * BLOCK_K + 4 * (j & 15) + i;
                dst[u][i] = (j < JOBS && (seg_s || seg_o) && t < p.g_rows) ? __ldg(idx + t) : 0;
              }
            }
          };
          if (kb0 < kb1) load_idx(kb0, nx);
          for (int kb = kb0; kb < kb1; ++kb) {
            int id[PER][4];
#pragma unroll
            for (int u = 0; u < PER; ++u)
#pragma unroll
              for (int i = 0; i < 4; ++i) id[u][i] = nx[u][i];
            if (kb + 1 < kb1) load_idx(kb + 1, nx);
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&bars->full[stage]);
            const uint32_t b_dst = sB + stage * B_STAGE;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
              const int j = gl + 32 * u;
              if (j < JOBS) {
                const int n0 = nt * BN + (j >> 4) * 64;
                const uint32_t dst = b_dst + (j >> 4) * 8192 + (j & 15) * 512;
                if (n0 < p.g_din) tma_gather4(dst, &tmB, fb, n0, id[u][0], id[u][1], id[u][2], id[u][3]);
                else if (n0 >= p.g_din + p.g_dp)
                  tma_gather4(dst, &tmB, fb, n0 - p.g_din - p.g_dp, id[u][0], id[u][1], id[u][2], id[u][3]);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
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
    // ================================================================== epilogue (warps 2-5)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    int it = 0;
    if (!MN && p.tma_epi) {
      // bf16 output: registers -> swizzled smem -> TMA store, 64 columns at a time, this warp's 32 rows only
      const uint32_t sC = base + plan.c + q * p.cbufs * CHUNK_BYTES;
      const uint32_t sX = base + plan.aux + q * 2 * CHUNK_BYTES;
      float* sbias = reinterpret_cast<float*>(base_ptr + plan.bias) + q * BN;
      const uint32_t sw = (uint32_t)(lane & 7) << 4;           // 128B swizzle: 16-byte chunk j of row r sits at j ^ (r & 7)
      const uint32_t row_off = (uint32_t)lane * 128;
      uint32_t gc = 0;                                          // chunks stored so far (staging buffer parity)
      uint32_t xc = 0;                                          // aux chunks consumed so far
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        int mt, nt, sp;
        decode(tile, mt, nt, sp);
        const int as = it & 1, aphase = (it >> 1) & 1;
        const int row0 = mt * BLOCK_M + q * 32;
        const int row = row0 + lane;
        const int ncols = min(BN, p.N - nt * BN);
        const int nchunks = (ncols + CHUNK_N - 1) / CHUNK_N;
        const bool warp_rows = row0 < p.M;
        if (has_aux && lane == 0 && warp_rows) {
          for (int c = 0; c < min(2, nchunks); ++c) {
            const uint32_t xb = smem_u32(&bars->aux_full[q][(xc + c) & 1]);
            mbar_arrive_expect_tx(xb, CHUNK_BYTES);
            tma_load_2d(sX + ((xc + c) & 1) * CHUNK_BYTES, &tmX, xb, nt * BN + c * CHUNK_N, row0);
          }
        }
        if (p.bias) {
          for (int j = lane; j < BN; j += 32) sbias[j] = (nt * BN + j < p.N) ? __ldg(p.bias + nt * BN + j) : 0.f;
        }
        const float rs = (p.rowscale && row < p.M) ? __ldg(p.rowscale + row) : 1.f;
        __syncwarp();
        mbar_wait(smem_u32(&bars->tmem_full[as]), aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + c * CHUNK_N, r0);
          tmem_ld32(t_row + c * CHUNK_N + 32, r1);
          const uint32_t cbuf = sC + (p.cbufs == 2 ? (gc & 1) : 0) * CHUNK_BYTES;
          if (warp_rows) {
            // the TMA store that last read this staging buffer must have drained it
            if (lane == 0) { if (p.cbufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
            if (has_aux) mbar_wait(smem_u32(&bars->aux_full[q][xc & 1]), (xc >> 1) & 1);
          }
          __syncwarp();
          tmem_ld_wait();
          if (c == nchunks - 1) {
            // accumulator fully read: hand the TMEM stage back before the math / stores of the last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[as]));
          }
          if (warp_rows) {
            const uint32_t xbuf = sX + (xc & 1) * CHUNK_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(h ? r1[j] : r0[j]);
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 b = *reinterpret_cast<const float4*>(sbias + c * CHUNK_N + h * 32 + j);
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
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const uint32_t off = row_off + ((((uint32_t)(h * 4 + j4)) << 4) ^ sw);
                if (has_aux) {
                  uint4 a;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(xbuf + off));
                  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (!bf16_pos(w[e] & 0xFFFFu)) v[j4 * 8 + e * 2] = 0.f;
                    if (!bf16_pos(w[e] >> 16)) v[j4 * 8 + e * 2 + 1] = 0.f;
                  }
                }
                const uint32_t u0 = pack_bf16(v[j4 * 8], v[j4 * 8 + 1]), u1 = pack_bf16(v[j4 * 8 + 2], v[j4 * 8 + 3]);
                const uint32_t u2 = pack_bf16(v[j4 * 8 + 4], v[j4 * 8 + 5]), u3 = pack_bf16(v[j4 * 8 + 6], v[j4 * 8 + 7]);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(cbuf + off), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
              }
            }
            fence_proxy_async();                   // generic-proxy smem writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, cbuf, nt * BN + c * CHUNK_N, row0);
              bulk_commit();
              if (has_aux && c + 2 < nchunks) {    // every lane has read aux buffer (xc & 1): refill it
                const uint32_t xb = smem_u32(&bars->aux_full[q][xc & 1]);
                mbar_arrive_expect_tx(xb, CHUNK_BYTES);
                tma_load_2d(xbuf, &tmX, xb, nt * BN + (c + 2) * CHUNK_N, row0);
              }
            }
            ++gc;
            if (has_aux) ++xc;
          }
        }
      }
      if (lane == 0) bulk_wait<0>();
    } else {
      // fp32 output (split-K partials, small fp32 results): direct 16-byte stores, 32 columns at a time
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
                    if (!bf16_pos(w[e] & 0xFFFFu)) v[j4 * 8 + e * 2] = 0.f;
                    if (!bf16_pos(w[e] >> 16)) v[j4 * 8 + e * 2 + 1] = 0.f;
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
                uint4 u;
                u.x = pack_bf16(v[j], v[j + 1]); u.y = pack_bf16(v[j + 2], v[j + 3]);
                u.z = pack_bf16(v[j + 4], v[j + 5]); u.w = pack_bf16(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(dst + j) = u;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[as]));
      }
    }
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

struct Maps {
  CUtensorMap a, b, p, c, x;
};

template <int BN, bool MN, int GATHER>
int launch(const Maps& m, const TcParams& p, cudaStream_t stream) {
  const size_t smem = smem_plan(BN, p.stages, p.cbufs, p.tma_epi != 0, p.mask_aux != nullptr).total;
  CSG_REQUIRE(smem <= SMEM_LIMIT, "gemm_tc: shared-memory plan of %zu bytes exceeds the limit", smem);
  static bool configured = false;
  if (!configured) {
    CSG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    configured = true;
  }
  int tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = tiles < csg_num_sms() ? tiles : csg_num_sms();
  gemm_tc_kernel<BN, MN, GATHER><<<grid, NUM_THREADS + (GATHER != G_NONE ? GATHER_WARPS * 32 : 0), smem, stream>>>(m.a, m.b, m.p, m.c, m.x, p);
  CSG_CHECK_LAUNCH("csg_gemm_bf16");
  return 0;
}

template <bool MN, int GATHER>
int launch_bn(int BN, const Maps& m, const TcParams& p, cudaStream_t stream) {
  switch (BN) {
    case 64: return launch<64, MN, GATHER>(m, p, stream);
    case 128: return launch<128, MN, GATHER>(m, p, stream);
    case 192: return launch<192, MN, GATHER>(m, p, stream);
    case 256: return launch<256, MN, GATHER>(m, p, stream);
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

// split-K plan of the MN-major (weight-gradient) GEMMs: one work item per CTA (a second wave would double the time)
void mn_split_plan(int M, int N, int K, int BN, int& kb_per_split, int& nsplits) {
  const int tiles = csg_div_up(M, BLOCK_M) * csg_div_up(N, BN);
  const int kb_total = csg_div_up(K, BLOCK_K);
  int splits = csg_num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  if (splits > kb_total) splits = kb_total;
  kb_per_split = csg_div_up(kb_total, splits);
  nsplits = csg_div_up(kb_total, kb_per_split);
}

int mn_pick_bn(int N, int gather) {
  if (gather == 2) return (N % 192 == 0) ? 192 : (N % 128 == 0 ? 128 : 64);
  return pick_bn(N);
}

}  // namespace

// bytes of split-K partials for an MN-major GEMM of this shape on the current device (gathered-B GEMMs may pick a
// narrower tile: the larger of the two plans is returned)
CSG_API size_t csg_gemm_bf16_workspace(int M, int N, int K, int mn_major) {
  if (!mn_major || M <= 0 || N <= 0 || K <= 0) return 0;
  size_t need = 0;
  for (int gather = 0; gather <= 2; gather += 2) {
    int kbs, ns;
    mn_split_plan(M, N, K, mn_pick_bn(N, gather), kbs, ns);
    const size_t b = ns > 1 ? (size_t)ns * M * N * sizeof(float) : 0;
    if (b > need) need = b;
  }
  return need;
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
                          int g_din, int g_dp, int g_ldp, int g_nobj,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (M == 0 || N == 0) return 0;
  if (K == 0 && mn_major && out_f32 && M > 0 && N > 0) {      // empty reduction (no triples): the gradient is zero
    CSG_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, stream));
    return 0;
  }
  CSG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad sizes M=%d N=%d K=%d", M, N, K);
  CsgProfScope prof(CSG_PROF_GEMM_BF16, 2.0 * M * N * K, stream);
  CSG_REQUIRE(N % 32 == 0, "gemm_bf16: N=%d must be a multiple of 32", N);
  CSG_REQUIRE((ldc % 8) == 0 || out_f32, "gemm_bf16: bf16 ldc must be a multiple of 8");
  CSG_REQUIRE(!out_f32 || (ldc % 4) == 0, "gemm_bf16: fp32 ldc must be a multiple of 4");
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.out_f32 = out_f32;
  p.bias = bias; p.relu = relu; p.rowscale = rowscale;
  p.mask_aux = reinterpret_cast<const __nv_bfloat16*>(mask_aux); p.ld_aux = ld_aux;
  p.g_sidx = g_sidx; p.g_oidx = g_oidx; p.g_din = g_din; p.g_dp = g_dp;
  p.g_rows = mn_major ? K : M;
  if (gather) {
    CSG_REQUIRE(g_obj && g_pred && g_sidx && g_oidx && g_nobj > 0, "gemm_bf16: gather sources missing");
    CSG_REQUIRE(g_din % 64 == 0 && g_dp % 64 == 0 && g_ldp % 8 == 0, "gemm_bf16: gather dims must be multiples of 64");
    CSG_REQUIRE((gather == 1 && !mn_major && K == 2 * g_din + g_dp) || (gather == 2 && mn_major && N == 2 * g_din + g_dp),
                "gemm_bf16: gather mode / shape mismatch");
  }
  if (p.mask_aux) CSG_REQUIRE(ld_aux % 8 == 0, "gemm_bf16: ld_aux must be a multiple of 8");
  // a gathered 64-column chunk must not straddle two source segments: guaranteed by %64 dims
  const int BN = mn_pick_bn(N, gather);
  p.m_tiles = csg_div_up(M, BLOCK_M);
  p.n_tiles = csg_div_up(N, BN);
  p.kb_total = csg_div_up(K, BLOCK_K);
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  void* out = C;
  p.tma_epi = 0;
  p.stages = 4;
  p.cbufs = 1;
  if (!mn_major) {
    CSG_REQUIRE(K % 8 == 0, "gemm_bf16: K=%d must be a multiple of 8", K);
    if (gather == 1) {
      // A = [obj[s] | pred | obj[o]]: object table rows by tile::gather4 (box {64, 1}), predicate rows by tile loads
      if (int rc = make_map(&maps.a, g_obj, g_din, g_nobj, (uint64_t)g_din * 2, 1)) return rc;
      if (int rc = make_map(&maps.p, g_pred, g_dp, M, (uint64_t)g_ldp * 2, BLOCK_M)) return rc;
    } else {
      if (int rc = make_map(&maps.a, A, K, M, (uint64_t)lda * 2, BLOCK_M)) return rc;
    }
    if (int rc = make_map(&maps.b, B, K, N, (uint64_t)ldb * 2, BN)) return rc;
    p.tma_epi = (!out_f32 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0;
    if (p.tma_epi) {
      if (int rc = make_map(&maps.c, C, N, M, (uint64_t)ldc * 2, 32)) return rc;
      if (p.mask_aux) { if (int rc = make_map(&maps.x, mask_aux, N, M, (uint64_t)ld_aux * 2, 32)) return rc; }
    }
    // deepest ring that fits, then double-buffered staging if it still fits
    const bool aux = p.mask_aux != nullptr;
    int best_s = 0, best_c = 1;
    for (int st = 4; st >= 2 && !best_s; --st)
      for (int cb = 2; cb >= 1; --cb)
        if (smem_plan(BN, st, cb, p.tma_epi != 0, aux).total <= SMEM_LIMIT) { best_s = st; best_c = cb; break; }
    CSG_REQUIRE(best_s > 0, "gemm_bf16: no shared-memory plan fits BN=%d", BN);
    p.stages = best_s; p.cbufs = best_c;
  } else {
    CSG_REQUIRE(out_f32, "gemm_bf16: MN-major (weight-gradient) GEMMs write fp32");
    CSG_REQUIRE(!bias && !relu && !rowscale && !mask_aux, "gemm_bf16: MN-major GEMMs have no epilogue");
    mn_split_plan(M, N, K, BN, p.kb_per_split, p.splits);
    if (p.splits > 1) {
      CSG_REQUIRE(ldc == N, "gemm_bf16: split-K output must be contiguous");
      CSG_REQUIRE(workspace && workspace_bytes >= (size_t)p.splits * M * N * sizeof(float), "gemm_bf16: workspace too small");
      p.C = workspace;
    }
    if (int rc = make_map(&maps.a, A, M, K, (uint64_t)lda * 2, 64)) return rc;
    if (gather == 2) {
      if (int rc = make_map(&maps.b, g_obj, g_din, g_nobj, (uint64_t)g_din * 2, 1)) return rc;
      if (int rc = make_map(&maps.p, g_pred, g_dp, K, (uint64_t)g_ldp * 2, 64)) return rc;
    } else {
      if (int rc = make_map(&maps.b, B, N, K, (uint64_t)ldb * 2, 64)) return rc;
    }
    p.stages = smem_plan(BN, 5, 1, false, false).total <= SMEM_LIMIT ? 5 : 4;
  }
  int rc;
  if (!mn_major) rc = gather == 1 ? launch_bn<false, G_A>(BN, maps, p, stream) : launch_bn<false, G_NONE>(BN, maps, p, stream);
  else rc = gather == 2 ? launch_bn<true, G_B>(BN, maps, p, stream) : launch_bn<true, G_NONE>(BN, maps, p, stream);
  if (rc) return rc;
  if (p.splits > 1) {
    long long MN = (long long)M * N;
    splitk_reduce_tc_kernel<<<csg_div_up(MN, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(p.C),
                                                                     reinterpret_cast<float*>(out), MN, p.splits);
    CSG_CHECK_LAUNCH("csg_gemm_bf16 split-K reduce");
  }
  return 0;
}
