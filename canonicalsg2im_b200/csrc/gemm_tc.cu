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
// Kernel anatomy (persistent, warp specialised, 320 threads; 448 in the gather variants):
//   warp 0      producer: TMA tile loads (lane 0) and, in the gather variants, one TMA tile::gather4 per lane
//               (4 gathered rows of 128 bytes each, written in the same 128-byte swizzle) into the smem ring
//   warp 1      MMA issuer (one lane)              tcgen05.mma + tcgen05.commit -> empty / tmem_full
//   warps 2-9   epilogue: tcgen05.ld 32x32b.x32 -> bias / ReLU / row scale -> bf16 -> swizzled smem staging ->
//               coalesced 16-byte global stores, with the ReLU mask applied in that coalesced layout (each warp owns
//               its 32 rows and its private bias row: no barrier spans more than one warp)
//   warps 10-13 (gather variants) TMA tile::gather4 issue, 8 lanes each
// TMEM: 512 columns = 2 accumulator stages x 256, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "internal.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int GATHER_WARPS = 4;               // extra warps of the gather variants, behind the epilogue warps (8 issuing lanes each)
constexpr int ACC_STRIDE = 256;                   // TMEM columns per accumulator stage
// Epilogue warps (template parameter EPW, 8 in every shipped instantiation): EPW / 4 per TMEM lane quarter.  The staged
// bf16 epilogue comes in two geometries (template flag SUB32): 64-column chunks (stores of full 128-byte lines) and
// 32-column sub-blocks (half the registers per warp, one tcgen05.ld per step); csg_gemm_bf16_deferred picks per shape.
constexpr int SUB_N = 32;                         // epilogue sub-block (SUB32): 32 columns = one tcgen05.ld.32x32b.x32 per lane
constexpr int SUB_BYTES = 32 * SUB_N * 2;         // per-warp staging buffer: 32 rows x 64 bytes
constexpr int CHUNK_N = 64;                       // epilogue chunk (!SUB32): 64 bf16 columns = one 128-byte swizzle row
constexpr int CHUNK_BYTES = 32 * CHUNK_N * 2;     // per-warp staging buffer: 32 rows x 128 bytes
constexpr size_t SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA

enum { G_NONE = 0, G_A = 1, G_B = 2 };

struct TcParams {
  int M, N, K;                 // K = reduction length (rows of the MN-major operands)
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int stages;                  // smem ring depth
  int smem_epi;                // bf16 output through the swizzled smem staging tile + coalesced stores (K-major kernels)
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
  const int* g_pidx;           // non-NULL: the predicate segment is gathered too (tmP = a table, row g_pidx[t]): layer 0
                               // reads its rows straight from the embedding tables (sg2im/model.py:108-109)
  int g_din, g_dp, g_rows;
  int debug;                   // scratch/bench_gemm.py only: 1 = no operand loads, 2 = no MMAs, 4 = no epilogue work
  int a_f16, b_f16, c_f16;     // operand / 16-bit output element formats: 0 = bf16, 1 = fp16 (kind::f16 takes both, per operand)
  // Tail split (K-major kernels with the staged epilogue): when the last round of the static tile schedule would leave
  // more than half of the workers idle, its tiles (indices >= tail_from) are cut into two half-width tiles each, so the
  // round costs half a tile time (dhid: 918 pair tiles on 74 pairs = 12.4 rounds -> 12.5 instead of 13).  -1: off.
  int tail_from, num_items;
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
// Bounded wait: a protocol bug traps instead of hanging the GPU box.  The report lives in one out-of-line function so
// that the many wait sites stay small (instruction cache) and keep no printf argument buffers on their stacks.
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("csg gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) mbar_timeout(bar, parity);
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
// cta_group::2 forms: the destination is this CTA's shared memory, the mbarrier lives in the pair's leader CTA
// (bar is a shared::cluster address obtained with mapa).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

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
// cta_group::2 forms (CTA pair: one MMA of M = 256 spans both SMs, each CTA holds half of B)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this CTA-relative address in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
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
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A format [7,10), B format [10,13) (kind::f16:
// 0 = fp16, 1 = bf16, chosen per operand), a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23),
// M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn, bool a_f16 = false, bool b_f16 = false) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// Shared-memory plan (offsets from the 1024-byte aligned base):
//   [A ring: stages x 16 KB][B ring: stages x BN*128 B][C staging: 4 warps x 4 KB]
//   [aux staging: 4 warps x 2 x 4 KB][bias: 4 warps x BN floats][barriers]
struct SmemPlan {
  uint32_t b, c, bias, bars, total;
};
__host__ __device__ inline SmemPlan smem_plan(int BN, int b_rows, int stages, bool smem_epi, int mt, int epw, bool sub32) {
  SmemPlan s;
  (void)BN;
  s.b = (uint32_t)stages * mt * A_STAGE_BYTES;
  s.c = s.b + (uint32_t)stages * b_rows * 128;
  s.bias = s.c + (smem_epi ? epw * (sub32 ? SUB_BYTES : CHUNK_BYTES) : 0);
  s.bars = s.bias + (smem_epi ? epw * (sub32 ? SUB_N : CHUNK_N) * 4 : 0);        // bias of the sub-block a warp is working on (one private row per warp)
  s.total = s.bars + (uint32_t)sizeof(Barriers);
  return s;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 16-bit output pair in the requested format; fp16 saturates at +-65504 instead of producing inf
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) {
  if (!f16) return pack_bf16(a, b);
  __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
// bf16 > 0  <=>  sign bit clear and magnitude non-zero
__device__ __forceinline__ bool bf16_pos(uint32_t h) { return h != 0 && h < 0x8000u; }

// tmA/tmB: operand maps (in the gather variants the gathered operand's slot holds the object-row table,
// box {64, 1}); tmP: predicate rows of the fused gather.
//
// CG = 2 (K-major variants only): the kernel runs as clusters of two CTAs on one TPC.  A pair owns a 256 x BN tile:
// each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader (cluster rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B across the pair, and each CTA's TMEM receives the
// accumulators of its own 128 rows.  This halves the B bytes every SM has to ingest per MMA cycle, which is what
// bounds the 1-CTA kernel (ncu r01d: 47 % tensor-pipe active with L2 and DRAM far from saturated).
// Pair protocol: all TMA loads of both CTAs complete on the LEADER's full[] barrier; the leader's commits are
// multicast to empty[] / tmem_full[] of both CTAs; both CTAs' epilogue warps arrive on the leader's tmem_empty[].
//
// MT > 1 (gathered-B weight gradient dW1 only): one CTA owns MT 128-row m sub-tiles of the output for one 128-column
// n-tile, i.e. MT accumulators side by side in TMEM.  The gathered operand (the TMA gather4 issue rate is what bounds
// this GEMM) is then staged once per k-block for MT * 128 rows of M instead of once per 128-row m-tile.
template <int BN, bool MN, int GATHER, int CG, int MT, int EPW, bool SUB32>
__global__ void __launch_bounds__(64 + EPW * 32 + (GATHER != G_NONE ? GATHER_WARPS * 32 : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmP, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];             // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* base_ptr = smem_raw;
  if ((base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("csg gemm_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  static_assert(CG == 1 || !MN, "CTA pairs are implemented for the K-major kernels");
  static_assert(MT == 1 || (MN && GATHER == G_B && CG == 1 && MT * BN <= 512), "MT > 1: gathered-B MN-major kernels only");
  static_assert(EPW == 8 || EPW == 16, "epilogue warps: 2 or 4 per TMEM lane quarter");
  constexpr int EPI_WARPS = EPW;
  constexpr int FIRST_GATHER_WARP = 2 + EPW;
  constexpr int NSUB = EPW / 4;                      // epilogue warps per TMEM lane quarter
  constexpr int A_STAGE = MT * A_STAGE_BYTES;        // bytes of the A tile(s) of one stage
  constexpr int B_ROWS = BN / CG;                   // rows of the B tile staged by this CTA
  constexpr int B_STAGE = B_ROWS * BLOCK_K * 2;
  constexpr int TILE_M = BLOCK_M * CG * MT;         // rows of C per work item (pair tile when CG = 2)
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int worker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nworkers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool has_aux = GATHER == G_NONE && p.mask_aux != nullptr;   // the gathered GEMMs never take a ReLU mask
  const SmemPlan plan = smem_plan(BN, B_ROWS, p.stages, p.smem_epi != 0, MT, EPW, SUB32);
  const uint32_t sA = base, sB = base + plan.b;
  Barriers* bars = reinterpret_cast<Barriers*>(base_ptr + plan.bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_items;
  const int STAGES = p.stages;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), EPI_WARPS * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (GATHER != G_NONE) tma_prefetch_desc(&tmP);
  }
  if (warp == 2) {
    if (CG == 2) { tmem_alloc_pair(smem_u32(&bars->tmem_base), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();     // barriers of BOTH CTAs are initialised past this point
  tc_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global memory and may run while the previous kernel of the stream drains; its results are visible past this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = bars->tmem_base;
  // shared::cluster address of a barrier of the pair's leader (identity for the 1-CTA kernel)
  auto leader = [&](const uint64_t* bar) -> uint32_t { return CG == 2 ? mapa(smem_u32(bar), 0u) : smem_u32(bar); };

  // work item -> (m-tile, n-tile, split) and the columns [n0, n0 + nw) of C it covers (a half tile of the split tail
  // covers BN / 2 of them)
  auto decode = [&](int tile, int& mt, int& nt, int& sp, int& n0, int& nw) {
    int half = -1;
    if (p.tail_from >= 0 && tile >= p.tail_from) {
      const int h = tile - p.tail_from;
      half = h & 1;
      tile = p.tail_from + (h >> 1);
    }
    nt = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    mt = r % p.m_tiles;
    sp = r / p.m_tiles;
    n0 = nt * BN;
    nw = BN;
    if (half >= 0) { nw = BN / 2; n0 += half * nw; }
  };
  auto kb_range = [&](int sp, int& kb0, int& kb1) {
    kb0 = sp * p.kb_per_split;
    kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    // ================================================================== producer (one lane): tile loads
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = worker; tile < num_tiles; tile += nworkers) {
        int mt, nt, sp, kb0, kb1, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        kb_range(sp, kb0, kb1);
        const int m0 = mt * TILE_M + (int)cta_rank * BLOCK_M;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fb = leader(&bars->full[stage]);
          // the whole stage (of both CTAs of a pair) is accounted here by the leader; bytes gathered by the gather warps or
          // loaded by the peer CTA may land before or after this arrive
          if (p.debug & 1) {
            if (CG == 1 || cta_rank == 0) mbar_arrive(smem_u32(&bars->full[stage]));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (CG == 1 || cta_rank == 0) mbar_arrive_expect_tx(smem_u32(&bars->full[stage]), CG * (A_STAGE + B_STAGE));
          const uint32_t a_dst = sA + stage * A_STAGE, b_dst = sB + stage * B_STAGE;
          if (!MN && CG == 2) {
            if (GATHER == G_NONE) tma_load_2d_pair(a_dst, &tmA, fb, kb * BLOCK_K, m0);
            else {
              const int c0 = kb * BLOCK_K;
              if (!p.g_pidx && c0 >= p.g_din && c0 < p.g_din + p.g_dp) tma_load_2d_pair(a_dst, &tmP, fb, c0 - p.g_din, m0);
            }
            // (a half tile loads the whole box as well: the MMA reads only its first nw / 2 rows of each CTA)
            tma_load_2d_pair(b_dst, &tmB, fb, kb * BLOCK_K, n0 + (int)cta_rank * (nw / 2));
          } else if (!MN) {
            // K-major: rows = M (or N), 64 contiguous k elements per 128-byte row
            if (GATHER == G_NONE) tma_load_2d(a_dst, &tmA, fb, kb * BLOCK_K, mt * BLOCK_M);
            else {
              const int c0 = kb * BLOCK_K;     // predicate block of the virtual row [obj[s] | pred | obj[o]]
              if (!p.g_pidx && c0 >= p.g_din && c0 < p.g_din + p.g_dp) tma_load_2d(a_dst, &tmP, fb, c0 - p.g_din, mt * BLOCK_M);
            }
            tma_load_2d(b_dst, &tmB, fb, kb * BLOCK_K, n0);
          } else {
            // MN-major: rows = k (64 per block), one [64 k x 64 mn] box per 64-wide chunk
            for (int c = 0; c < MT * BLOCK_M / 64; ++c) tma_load_2d(a_dst + c * 8192, &tmA, fb, mt * TILE_M + c * 64, kb * BLOCK_K);
            for (int c = 0; c < BN / 64; ++c) {
              const int n0 = nt * BN + c * 64;
              if (GATHER == G_NONE) tma_load_2d(b_dst + c * 8192, &tmB, fb, n0, kb * BLOCK_K);
              else if (!p.g_pidx && n0 >= p.g_din && n0 < p.g_din + p.g_dp) tma_load_2d(b_dst + c * 8192, &tmP, fb, n0 - p.g_din, kb * BLOCK_K);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= FIRST_GATHER_WARP) {
    // ================================================================== gather producers (warps 10-13, lanes 0-7)
    // Object rows of the virtual operand arrive by TMA tile::gather4 (4 rows x 128 bytes per instruction, written
    // in the same 128-byte swizzle as a tile load).  A TMA instruction takes warp-uniform operands, so a warp
    // issues them one lane at a time: the work is spread over 4 warps x 8 lanes.
    if (GATHER != G_NONE && lane < 8) {
      const int gl = (warp - FIRST_GATHER_WARP) * 8 + lane;          // 0..31
      int stage = 0, phase = 0;
      for (int tile = worker; tile < num_tiles; tile += nworkers) {
        int mt, nt, sp, kb0, kb1, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        kb_range(sp, kb0, kb1);
        if (GATHER == G_A) {
          // A tile [128 triples x 64 k]: this lane owns rows 4*gl .. 4*gl+3 (tmA = object table)
          int si[4], oi[4], pi[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int t = mt * TILE_M + (int)cta_rank * BLOCK_M + 4 * gl + i;
            const bool ok = t < p.g_rows;
            si[i] = ok ? __ldg(p.g_sidx + t) : 0;
            oi[i] = ok ? __ldg(p.g_oidx + t) : 0;
            pi[i] = (ok && p.g_pidx) ? __ldg(p.g_pidx + t) : 0;
          }
          for (int kb = kb0; kb < kb1; ++kb) {
            const int c0 = kb * BLOCK_K;
            const bool seg_s = c0 < p.g_din, seg_o = c0 >= p.g_din + p.g_dp;
            if (!seg_s && !seg_o && p.g_pidx) {
              // predicate segment straight from the predicate table (tmP: box {64, 1})
              mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
              const uint32_t fb = leader(&bars->full[stage]);
              const uint32_t dst = sA + stage * A_STAGE + gl * 512;
              if (CG == 2) tma_gather4_pair(dst, &tmP, fb, c0 - p.g_din, pi[0], pi[1], pi[2], pi[3]);
              else tma_gather4(dst, &tmP, fb, c0 - p.g_din, pi[0], pi[1], pi[2], pi[3]);
            } else if (seg_s || seg_o) {
              mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
              const uint32_t fb = leader(&bars->full[stage]);
              const uint32_t dst = sA + stage * A_STAGE + gl * 512;
              if (CG == 2) {
                if (seg_s) tma_gather4_pair(dst, &tmA, fb, c0, si[0], si[1], si[2], si[3]);
                else tma_gather4_pair(dst, &tmA, fb, c0 - p.g_din - p.g_dp, oi[0], oi[1], oi[2], oi[3]);
              } else {
                if (seg_s) tma_gather4(dst, &tmA, fb, c0, si[0], si[1], si[2], si[3]);
                else tma_gather4(dst, &tmA, fb, c0 - p.g_din - p.g_dp, oi[0], oi[1], oi[2], oi[3]);
              }
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
              const bool seg_p = !seg_s && !seg_o && p.g_pidx != nullptr;
              const int* idx = seg_s ? p.g_sidx : (seg_o ? p.g_oidx : p.g_pidx);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int t = kb * BLOCK_K + 4 * (j & 15) + i;
                dst[u][i] = (j < JOBS && (seg_s || seg_o || seg_p) && t < p.g_rows) ? __ldg(idx + t) : 0;
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
                else if (p.g_pidx)
                  tma_gather4(dst, &tmP, fb, n0 - p.g_din, id[u][0], id[u][1], id[u][2], id[u][3]);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0 && cta_rank == 0) {
      int stage = 0, phase = 0, it = 0;
      for (int tile = worker; tile < num_tiles; tile += nworkers, ++it) {
        int mt, nt, sp, kb0, kb1, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        kb_range(sp, kb0, kb1);
        const int as = it & 1, aphase = (it >> 1) & 1;
        // the last n-tile may be narrower than BN: issue MMAs of its real width (N % 32 == 0 is required by the host
        // side; an MMA of N = 192 costs as much as N = 256, which is why BN is 256 with a narrow tail and not 192)
        const int mma_n = !MN ? (CG == 1 ? min(nw, p.N - n0) : nw) : BN;
        const uint32_t idesc = make_idesc(BLOCK_M * CG, mma_n, MN, MN, p.a_f16 != 0, p.b_f16 != 0);
        mbar_wait(smem_u32(&bars->tmem_empty[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint32_t a_base = sA + stage * A_STAGE, b_base = sB + stage * B_STAGE;
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              uint64_t ad, bd;
              if (!MN) {
                ad = make_desc(a_base + k * 32, 16, 1024);
                bd = make_desc(b_base + k * 32, 16, 1024);
              } else {
                ad = make_desc(a_base + mi * A_STAGE_BYTES + k * 2048, 8192, 1024);
                bd = make_desc(b_base + k * 2048, 8192, 1024);
              }
              if (p.debug & 2) continue;
              if (CG == 2) umma_bf16_pair(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem + mi * BN, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          if (CG == 2) {
            umma_commit_pair(smem_u32(&bars->empty[stage]));   // frees the slot in both CTAs when these MMAs retire
            if (kb == kb1 - 1) umma_commit_pair(smem_u32(&bars->tmem_full[as]));
          } else {
            umma_commit(smem_u32(&bars->empty[stage]));        // frees the smem slot when these MMAs retire
            if (kb == kb1 - 1) umma_commit(smem_u32(&bars->tmem_full[as]));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0) {                                      // empty k range (never scheduled)
          if (CG == 2) umma_commit_pair(smem_u32(&bars->tmem_full[as])); else umma_commit(smem_u32(&bars->tmem_full[as]));
        }
      }
    }
  } else if (warp < FIRST_GATHER_WARP) {
    // ================================================================== epilogue (warps 2 .. 2+EPW-1)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int ew = warp - 2;                      // 0..EPW-1
    const int sub = ew >> 2;                      // the NSUB warps of a quarter take every NSUB-th 32-column sub-block
    int it = 0;
    if (!MN && p.smem_epi) {
      if constexpr (SUB32) {
      // bf16 output, 32 columns at a time: registers (one row per lane) -> bias / ReLU / row scale -> this warp's
      // swizzled 32 x 64 B staging tile -> registers in the coalesced layout (one instruction = 8 rows x 64 B) ->
      // ReLU mask (the mask operand is read with the same coalesced addressing, before the TMEM wait) -> 16-byte global
      // stores.  Stores are fire-and-forget: no epilogue warp ever waits for a write to drain, and no barrier spans
      // more than one warp (the bias of a sub-block sits in a private 128-byte row of the warp).
      const uint32_t sC = base + plan.c + ew * SUB_BYTES;
      float* sbias = reinterpret_cast<float*>(base_ptr + plan.bias) + ew * SUB_N;
      // staging rows are 64 bytes (4 pieces of 16); piece j of row r sits at j ^ ((r >> 1) & 3): the 8 lanes of a
      // shared-memory wavefront then touch 8 distinct 16-byte bank groups, for the row-per-lane writes and the
      // coalesced reads alike
      const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
      const uint32_t row_off = (uint32_t)lane * (SUB_N * 2);
      const int crow = lane >> 2, cpiece = lane & 3;           // coalesced layout: rows i*8 + crow, 16-byte piece cpiece
      __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(p.C);
      const bool cf16 = p.c_f16 != 0;
      for (int tile = worker; tile < num_tiles; tile += nworkers, ++it) {
        int mt, nt, sp, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        const int as = it & 1, aphase = (it >> 1) & 1;
        const int row0 = mt * TILE_M + (int)cta_rank * BLOCK_M + q * 32;
        const int row = row0 + lane;
        const int ncols = min(nw, p.N - n0);
        const int nsb = (ncols + SUB_N - 1) / SUB_N;
        const bool warp_rows = row0 < p.M;
        const float rs = (p.rowscale && row < p.M) ? __ldg(p.rowscale + row) : 1.f;
        mbar_wait(smem_u32(&bars->tmem_full[as]), aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
        const int my_last = nsb > sub ? ((nsb - 1 - sub) / NSUB) * NSUB + sub : -1;   // last sub-block of this warp
        if ((p.debug & 4) || my_last < 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(leader(&bars->tmem_empty[as])); else mbar_arrive(smem_u32(&bars->tmem_empty[as])); }
          continue;
        }
#pragma unroll 1
        for (int c = sub; c < nsb; c += NSUB) {
          const int col0 = n0 + c * SUB_N;                     // N % 32 == 0: a sub-block is never partial
          const float bv = p.bias ? __ldg(p.bias + col0 + lane) : 0.f;
          uint32_t r[32];
          tmem_ld32(t_row + c * SUB_N, r);
          // ReLU-mask operand of this sub-block, in the coalesced layout of the final stores: needed only after the math
          // and the staging round trip below, which cover its latency
          const int col = col0 + cpiece * 8;                   // first column of this lane's 16-byte piece
          uint4 ax[4];
          if (has_aux && warp_rows) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = row0 + i * 8 + crow;
              ax[i] = rr < p.M ? __ldg(reinterpret_cast<const uint4*>(p.mask_aux + (size_t)rr * p.ld_aux + col))
                               : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          tmem_ld_wait();
          if (c == my_last) {
            // accumulator fully read: hand the TMEM stage back before the math / stores of the last sub-block
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(leader(&bars->tmem_empty[as])); else mbar_arrive(smem_u32(&bars->tmem_empty[as])); }
          }
          if (warp_rows) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
              sbias[lane] = bv;
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + j);
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
              const uint32_t off = row_off + ((((uint32_t)j4) ^ sw) << 4);
              const uint32_t u0 = pack16(v[j4 * 8], v[j4 * 8 + 1], cf16), u1 = pack16(v[j4 * 8 + 2], v[j4 * 8 + 3], cf16);
              const uint32_t u2 = pack16(v[j4 * 8 + 4], v[j4 * 8 + 5], cf16), u3 = pack16(v[j4 * 8 + 6], v[j4 * 8 + 7], cf16);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sC + off), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = i * 8 + crow;
              uint4 o;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                           : "r"(sC + (uint32_t)rl * (SUB_N * 2) + ((uint32_t)(cpiece ^ ((rl >> 1) & 3)) << 4)));
              if (has_aux) {
                const uint32_t a[4] = {ax[i].x, ax[i].y, ax[i].z, ax[i].w};
                uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!bf16_pos(a[e] & 0xFFFFu)) w[e] &= 0xFFFF0000u;
                  if (!bf16_pos(a[e] >> 16)) w[e] &= 0x0000FFFFu;
                }
                o = make_uint4(w[0], w[1], w[2], w[3]);
              }
              if (row0 + rl < p.M)
                *reinterpret_cast<uint4*>(Cb + (size_t)(row0 + rl) * p.ldc + col) = o;
            }
            __syncwarp();                        // the staging tile and the bias row are rewritten by the next sub-block
          }
        }
      }
      } else {
      // bf16 output, 64 columns at a time: registers (one row per lane) -> bias / ReLU / row scale -> this warp's
      // swizzled 32 x 128 B staging tile -> registers in the coalesced layout (one instruction = 4 rows x 128 B) ->
      // ReLU mask (the mask operand is read with the same coalesced addressing, one chunk ahead) -> 16-byte global
      // stores.  Stores are fire-and-forget: no epilogue warp ever waits for a write to drain.
      const uint32_t sC = base + plan.c + ew * CHUNK_BYTES;
      float* sbias = reinterpret_cast<float*>(base_ptr + plan.bias) + ew * CHUNK_N;   // private bias row of this warp
      const uint32_t sw = (uint32_t)(lane & 7) << 4;           // 128B swizzle: 16-byte piece j of row r sits at j ^ (r & 7)
      const uint32_t row_off = (uint32_t)lane * 128;
      const int crow = lane >> 3, cpiece = lane & 7;           // coalesced layout: rows i*4 + crow, 16-byte piece cpiece
      __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(p.C);
      const bool cf16 = p.c_f16 != 0;
      auto load_mask = [&](uint4 (&ax)[8], int row0, int col) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + i * 4 + crow;
          ax[i] = (r < p.M && col < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.mask_aux + (size_t)r * p.ld_aux + col))
                                         : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      for (int tile = worker; tile < num_tiles; tile += nworkers, ++it) {
        int mt, nt, sp, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        const int as = it & 1, aphase = (it >> 1) & 1;
        const int row0 = mt * TILE_M + (int)cta_rank * BLOCK_M + q * 32;
        const int row = row0 + lane;
        const int ncols = min(nw, p.N - n0);
        const int nchunks = (ncols + CHUNK_N - 1) / CHUNK_N;
        const bool warp_rows = row0 < p.M;
        const float rs = (p.rowscale && row < p.M) ? __ldg(p.rowscale + row) : 1.f;
        mbar_wait(smem_u32(&bars->tmem_full[as]), aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
        const int my_last = nchunks > sub ? ((nchunks - 1 - sub) / NSUB) * NSUB + sub : -1;   // last chunk of this warp
        if ((p.debug & 4) || my_last < 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(leader(&bars->tmem_empty[as])); else mbar_arrive(smem_u32(&bars->tmem_empty[as])); }
          continue;
        }
#pragma unroll 1
        for (int c = sub; c < nchunks; c += NSUB) {
          const int bc = n0 + c * CHUNK_N + lane;
          const float bv0 = (p.bias && bc < p.N) ? __ldg(p.bias + bc) : 0.f;
          const float bv1 = (p.bias && bc + 32 < p.N) ? __ldg(p.bias + bc + 32) : 0.f;
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + c * CHUNK_N, r0);
          tmem_ld32(t_row + c * CHUNK_N + 32, r1);
          const int col = n0 + c * CHUNK_N + cpiece * 8;       // first column of this lane's 16-byte piece
          // ReLU-mask operand of this chunk, in the coalesced layout of the final stores: needed only after the math
          // and the staging round trip below, which cover its latency
          uint4 ax[8];
          if (has_aux && warp_rows) load_mask(ax, row0, col);
          tmem_ld_wait();
          if (c == my_last) {
            // accumulator fully read: hand the TMEM stage back before the math / stores of the last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(leader(&bars->tmem_empty[as])); else mbar_arrive(smem_u32(&bars->tmem_empty[as])); }
          }
          if (warp_rows) {
            if (p.bias) {
              sbias[lane] = bv0; sbias[lane + 32] = bv1;
              __syncwarp();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(h ? r1[j] : r0[j]);
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 b = *reinterpret_cast<const float4*>(sbias + h * 32 + j);
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
                const uint32_t u0 = pack16(v[j4 * 8], v[j4 * 8 + 1], cf16), u1 = pack16(v[j4 * 8 + 2], v[j4 * 8 + 3], cf16);
                const uint32_t u2 = pack16(v[j4 * 8 + 4], v[j4 * 8 + 5], cf16), u3 = pack16(v[j4 * 8 + 6], v[j4 * 8 + 7], cf16);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sC + off), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + crow;
              uint4 o;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                           : "r"(sC + (uint32_t)rl * 128 + ((uint32_t)(cpiece ^ (rl & 7)) << 4)));
              if (has_aux) {
                const uint32_t a[4] = {ax[i].x, ax[i].y, ax[i].z, ax[i].w};
                uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!bf16_pos(a[e] & 0xFFFFu)) w[e] &= 0xFFFF0000u;
                  if (!bf16_pos(a[e] >> 16)) w[e] &= 0x0000FFFFu;
                }
                o = make_uint4(w[0], w[1], w[2], w[3]);
              }
              if (row0 + rl < p.M && col < p.N)
                *reinterpret_cast<uint4*>(Cb + (size_t)(row0 + rl) * p.ldc + col) = o;
            }
            __syncwarp();                        // the staging tile and the bias row are rewritten by the next chunk
          }
        }
      }
      }
    } else {
      // fp32 output (split-K partials, small fp32 results): direct 16-byte stores, 32 columns at a time
      for (int tile = worker; tile < num_tiles; tile += nworkers, ++it) {
        int mt, nt, sp, n0, nw;
        decode(tile, mt, nt, sp, n0, nw);
        const int as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bars->tmem_full[as]), aphase);
        tc_fence_after();
#pragma unroll 1
        for (int ci = sub; ci < MT * (BN / 32); ci += NSUB) {
          const int mi = ci / (BN / 32), c = ci % (BN / 32);          // m sub-tile (MT > 1), 32-column group
          const int row = mt * TILE_M + (int)cta_rank * BLOCK_M + mi * BLOCK_M + q * 32 + lane;
          const bool rowok = row < p.M;
          const float rs = (p.rowscale && rowok) ? p.rowscale[row] : 1.f;
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE + mi * BN;
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
                const bool cf16 = p.c_f16 != 0;
                u.x = pack16(v[j], v[j + 1], cf16); u.y = pack16(v[j + 2], v[j + 3], cf16);
                u.z = pack16(v[j + 4], v[j + 5], cf16); u.w = pack16(v[j + 6], v[j + 7], cf16);
                *reinterpret_cast<uint4*>(dst + j) = u;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(leader(&bars->tmem_empty[as])); else mbar_arrive(smem_u32(&bars->tmem_empty[as])); }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // no CTA leaves while its peer can still touch its smem / TMEM
  if (warp == 2) { if (CG == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
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
  CUtensorMap a, b, p;
};

template <int BN, bool MN, int GATHER, int CG, int MT = 1, int EPW = 8, bool SUB32 = false>
int launch(const Maps& m, const TcParams& p, cudaStream_t stream) {
  const size_t smem = smem_plan(BN, BN / CG, p.stages, p.smem_epi != 0, MT, EPW, SUB32).total;
  CSG_REQUIRE(smem <= SMEM_LIMIT, "gemm_tc: shared-memory plan of %zu bytes exceeds the limit", smem);
  static bool configured = false;
  if (!configured) {
    CSG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN, GATHER, CG, MT, EPW, SUB32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  const int nthreads = 64 + EPW * 32 + (GATHER != G_NONE ? GATHER_WARPS * 32 : 0);
  if (CG == 1) {
    const int grid = tiles < csg_num_sms() ? tiles : csg_num_sms();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(nthreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CSG_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MN, GATHER, CG, MT, EPW, SUB32>, m.a, m.b, m.p, p));
  } else {
    // one CTA pair (cluster of 2 on one TPC) per work item, persistent over the pair tiles
    const int pairs = csg_num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    cfg.blockDim = dim3(nthreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    CSG_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MN, GATHER, CG, MT, EPW, SUB32>, m.a, m.b, m.p, p));
  }
  CSG_CHECK_LAUNCH("csg_gemm_bf16");
  return 0;
}

template <bool MN, int GATHER, int CG, int EPW = 8, bool SUB32 = false>
int launch_bn(int BN, const Maps& m, const TcParams& p, cudaStream_t stream) {
  switch (BN) {
    case 64: return launch<64, MN, GATHER, CG, 1, EPW, SUB32>(m, p, stream);
    case 128: return launch<128, MN, GATHER, CG, 1, EPW, SUB32>(m, p, stream);
    case 192: return launch<192, MN, GATHER, CG, 1, EPW, SUB32>(m, p, stream);
    case 256: return launch<256, MN, GATHER, CG, 1, EPW, SUB32>(m, p, stream);
  }
  csg_set_error("gemm_tc: unsupported BLOCK_N %d", BN);
  return 1;
}

// CTA-pair policy of the K-major GEMMs: -1 = automatic (pairs once M fills every SM at least twice and the reduction
// is long, K >= 1024: measured 132 vs 136 us on the dhid shape, while short-K shapes are 3-5 % faster on single CTAs),
// 0 = never, 1 = whenever the shape allows it (tests)
int g_pair_mode = -1;

// K-major kernels: 256-wide tiles with a narrower last tile
// (the MMA of the tail uses its real width); short M (the per-object GEMMs of net2 / box_net) takes narrower tiles until
// every SM has one
int kmajor_pick_bn(int M, int N) {
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  const int m_tiles = csg_div_up(M, BLOCK_M);
  while (bn > 64 && m_tiles * csg_div_up(N, bn) < csg_num_sms()) bn >>= 1;
  return bn;
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
void mn_split_plan(int M, int N, int K, int BN, int& kb_per_split, int& nsplits, int mt = 1) {
  const int tiles = csg_div_up(M, BLOCK_M * mt) * csg_div_up(N, BN);
  const int kb_total = csg_div_up(K, BLOCK_K);
  int splits = csg_num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  if (splits > kb_total) splits = kb_total;
  kb_per_split = csg_div_up(kb_total, splits);
  nsplits = csg_div_up(kb_total, kb_per_split);
}

// gathered-B weight gradient with 128-column n-tiles: GB_MT m sub-tiles (accumulators) per CTA, so that the gathered
// operand is staged once per k-block for GB_MT * 128 rows of M.  (GB_MT = 4 halves the gathers again but leaves room
// for only two 80 KB stages and measured no faster than GB_MT = 1; 2 keeps a 4-stage ring.)
constexpr int GB_MT = 2;
bool mn_gather_mt4(int M, int N, int gather) { return gather == 2 && (M % (GB_MT * BLOCK_M)) == 0 && (N % 128) == 0; }

int mn_pick_bn(int M, int N, int gather) {
  if (mn_gather_mt4(M, N, gather)) return 128;
  if (gather == 2) return (N % 192 == 0) ? 192 : (N % 128 == 0 ? 128 : 64);
  return pick_bn(N);
}

}  // namespace

CSG_API void csg_gemm_bf16_set_pair_mode(int mode) { g_pair_mode = mode; }

// bytes of split-K partials for an MN-major GEMM of this shape on the current device (gathered-B GEMMs may pick a
// narrower tile: the larger of the two plans is returned)
CSG_API size_t csg_gemm_bf16_workspace(int M, int N, int K, int mn_major) {
  if (!mn_major || M <= 0 || N <= 0 || K <= 0) return 0;
  size_t need = 0;
  for (int gather = 0; gather <= 2; gather += 2) {
    int kbs, ns;
    mn_split_plan(M, N, K, mn_pick_bn(M, N, gather), kbs, ns, mn_gather_mt4(M, N, gather) ? GB_MT : 1);
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
// out_f32 selects fp32 or 16-bit C.  bias [N] fp32, rowscale [M] fp32, mask_aux [M, ld_aux] (16-bit, only its sign is
// used) may be null.  formats: bit 0 = A (or the gathered rows of gather = 1) is fp16, bit 1 = B (or the gathered rows
// of gather = 2) is fp16, bit 2 = the 16-bit C is written as fp16 (saturating); clear bits mean bf16.
CSG_API int csg_gemm_bf16(int mn_major, int gather, int M, int N, int K,
                          const void* A, int lda, const void* B, int ldb, void* C, int ldc, int out_f32,
                          const float* bias, int relu, const float* rowscale, const void* mask_aux, int ld_aux,
                          const void* g_obj, const void* g_pred, const int* g_sidx, const int* g_oidx,
                          int g_din, int g_dp, int g_ldp, int g_nobj, const int* g_pidx, int g_npred, int formats,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  CsgReduceJob job;
  if (int rc = csg_gemm_bf16_deferred(mn_major, gather, M, N, K, A, lda, B, ldb, C, ldc, out_f32, bias, relu, rowscale,
                                      mask_aux, ld_aux, g_obj, g_pred, g_sidx, g_oidx, g_din, g_dp, g_ldp, g_nobj, g_pidx, g_npred,
                                      formats, workspace, workspace_bytes, stream, &job)) return rc;
  return job.parts > 0 ? csg_reduce_multi(&job, 1, stream) : 0;     // split-K final pass (fixed order: split 0, 1, ...)
}

int csg_gemm_bf16_deferred(int mn_major, int gather, int M, int N, int K,
                           const void* A, int lda, const void* B, int ldb, void* C, int ldc, int out_f32,
                           const float* bias, int relu, const float* rowscale, const void* mask_aux, int ld_aux,
                           const void* g_obj, const void* g_pred, const int* g_sidx, const int* g_oidx,
                           int g_din, int g_dp, int g_ldp, int g_nobj, const int* g_pidx, int g_npred, int formats,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream, CsgReduceJob* job) {
  job->parts = 0; job->n = 0; job->partial = nullptr; job->out = nullptr; job->stride = 0; job->lanes = 1;
  job->op = CSG_RED_SUM; job->aux = nullptr; job->ncols = 0; job->ldo = 0;
  if (M == 0 || N == 0) return 0;
  if (K == 0 && mn_major && out_f32 && M > 0 && N > 0) {      // empty reduction (no triples): the gradient is zero
    CSG_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, stream));
    return 0;
  }
  CSG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad sizes M=%d N=%d K=%d", M, N, K);
  CsgProfScope prof(CSG_PROF_GEMM_BF16, 2.0 * M * N * K, stream);   // the GEMM kernel itself (a deferred split-K pass is not in it)
  CSG_REQUIRE(N % 32 == 0, "gemm_bf16: N=%d must be a multiple of 32", N);
  CSG_REQUIRE((ldc % 8) == 0 || out_f32, "gemm_bf16: bf16 ldc must be a multiple of 8");
  CSG_REQUIRE(!out_f32 || (ldc % 4) == 0, "gemm_bf16: fp32 ldc must be a multiple of 4");
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.out_f32 = out_f32;
  p.bias = bias; p.relu = relu; p.rowscale = rowscale;
  p.mask_aux = reinterpret_cast<const __nv_bfloat16*>(mask_aux); p.ld_aux = ld_aux;
  p.g_sidx = g_sidx; p.g_oidx = g_oidx; p.g_din = g_din; p.g_dp = g_dp;
  p.g_pidx = gather ? g_pidx : nullptr;
  CSG_REQUIRE(!p.g_pidx || g_npred > 0, "gemm_bf16: g_pidx needs the row count of the predicate table");
  p.g_rows = mn_major ? K : M;
  p.a_f16 = formats & 1; p.b_f16 = (formats >> 1) & 1; p.c_f16 = (formats >> 2) & 1;
  { const char* dbg = getenv("CSG_GEMM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  if (gather) {
    CSG_REQUIRE(g_obj && g_pred && g_sidx && g_oidx && g_nobj > 0, "gemm_bf16: gather sources missing");
    CSG_REQUIRE(g_din % 64 == 0 && g_dp % 64 == 0 && g_ldp % 8 == 0, "gemm_bf16: gather dims must be multiples of 64");
    CSG_REQUIRE((gather == 1 && !mn_major && K == 2 * g_din + g_dp) || (gather == 2 && mn_major && N == 2 * g_din + g_dp),
                "gemm_bf16: gather mode / shape mismatch");
  }
  if (p.mask_aux) CSG_REQUIRE(ld_aux % 8 == 0, "gemm_bf16: ld_aux must be a multiple of 8");
  CSG_REQUIRE(!(gather && mask_aux), "gemm_bf16: the gathered GEMMs take no ReLU mask");
  // a gathered 64-column chunk must not straddle two source segments: guaranteed by %64 dims
  const int BN = mn_major ? mn_pick_bn(M, N, gather) : kmajor_pick_bn(M, N);
  const int MT = (mn_major && mn_gather_mt4(M, N, gather)) ? GB_MT : 1;
  // CTA pairs (cta_group::2): K-major only; each CTA stages BN/2 rows of B, so BN/2 must keep the 8-row swizzle atoms
  const bool pair = !mn_major && (BN % 32 == 0) && csg_num_sms() >= 2 &&
                    (g_pair_mode == 1 || (g_pair_mode < 0 && M >= 2 * BLOCK_M * csg_num_sms() && K >= 1024));
  const int CG = pair ? 2 : 1;
  // epilogue warps: 16 for the K-major kernels that stream their operands (F2, dhid, dX: their epilogue is a latency
  // chain that 8 warps do not hide), 8 for the gathered and the weight-gradient kernels (CSG_GEMM_EPW=8: everywhere)
  // epilogue geometry of the K-major kernels (measured A/B in one run, scratch/bench_gemm.py): 32-column sub-blocks are
  // 5 % faster on the gathered F1 (59.9 vs 62.8 us), 64-column chunks (full 128-byte lines per store) 2 % faster on F2;
  // 16 epilogue warps on sub-blocks cut the epilogue ALONE from 108 to 82 us but left the whole F2 slower (153 vs 144 us):
  // what bounds these kernels is the interference of loads, MMAs and epilogue on one SM, not the epilogue's own latency
  const int epw = 8;
  bool sub32 = !mn_major && gather == 1;
  { const char* e = getenv("CSG_GEMM_EPI");      // scratch/bench_gemm.py: "8x32" / "8x64" for every K-major kernel
    if (e && !mn_major) sub32 = !strcmp(e, "8x32"); }
  p.m_tiles = csg_div_up(M, BLOCK_M * CG * MT);
  p.n_tiles = csg_div_up(N, BN);
  p.kb_total = csg_div_up(K, BLOCK_K);
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  void* out = C;
  p.smem_epi = 0;
  p.stages = 4;
  if (!mn_major) {
    CSG_REQUIRE(K % 8 == 0, "gemm_bf16: K=%d must be a multiple of 8", K);
    if (gather == 1) {
      // A = [obj[s] | pred | obj[o]]: object table rows by tile::gather4 (box {64, 1}), predicate rows by tile loads
      if (int rc = make_map(&maps.a, g_obj, g_din, g_nobj, (uint64_t)g_din * 2, 1)) return rc;
      if (p.g_pidx) { if (int rc = make_map(&maps.p, g_pred, g_dp, g_npred, (uint64_t)g_ldp * 2, 1)) return rc; }
      else if (int rc = make_map(&maps.p, g_pred, g_dp, M, (uint64_t)g_ldp * 2, BLOCK_M)) return rc;
    } else {
      if (int rc = make_map(&maps.a, A, K, M, (uint64_t)lda * 2, BLOCK_M)) return rc;
    }
    if (int rc = make_map(&maps.b, B, K, N, (uint64_t)ldb * 2, BN / CG)) return rc;
    p.smem_epi = (!out_f32 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                  (!mask_aux || (reinterpret_cast<uintptr_t>(mask_aux) & 15) == 0)) ? 1 : 0;
    // deepest ring that fits
    int best_s = 0;
    for (int st = (pair ? 6 : 5); st >= 2 && !best_s; --st)
      if (smem_plan(BN, BN / CG, st, p.smem_epi != 0, 1, epw, sub32).total <= SMEM_LIMIT) best_s = st;
    { const char* e = getenv("CSG_GEMM_STAGES"); if (e && atoi(e) >= 3 && atoi(e) < best_s) best_s = atoi(e); }   // scratch/bench_gemm.py (3 stages: -4 %)
    CSG_REQUIRE(best_s > 0, "gemm_bf16: no shared-memory plan fits BN=%d", BN);
    p.stages = best_s;
  } else {
    CSG_REQUIRE(out_f32, "gemm_bf16: MN-major (weight-gradient) GEMMs write fp32");
    CSG_REQUIRE(!bias && !relu && !rowscale && !mask_aux, "gemm_bf16: MN-major GEMMs have no epilogue");
    mn_split_plan(M, N, K, BN, p.kb_per_split, p.splits, MT);
    if (p.splits > 1) {
      CSG_REQUIRE(ldc == N, "gemm_bf16: split-K output must be contiguous");
      CSG_REQUIRE(workspace && workspace_bytes >= (size_t)p.splits * M * N * sizeof(float), "gemm_bf16: workspace too small");
      p.C = workspace;
    }
    if (int rc = make_map(&maps.a, A, M, K, (uint64_t)lda * 2, 64)) return rc;
    if (gather == 2) {
      if (int rc = make_map(&maps.b, g_obj, g_din, g_nobj, (uint64_t)g_din * 2, 1)) return rc;
      if (p.g_pidx) { if (int rc = make_map(&maps.p, g_pred, g_dp, g_npred, (uint64_t)g_ldp * 2, 1)) return rc; }
      else if (int rc = make_map(&maps.p, g_pred, g_dp, K, (uint64_t)g_ldp * 2, 64)) return rc;
    } else {
      if (int rc = make_map(&maps.b, B, N, K, (uint64_t)ldb * 2, 64)) return rc;
    }
    p.stages = 5;
    while (p.stages > 2 && smem_plan(BN, BN, p.stages, false, MT, 8, false).total > SMEM_LIMIT) --p.stages;
    CSG_REQUIRE(MT == 1 || p.m_tiles * p.n_tiles * p.splits <= csg_num_sms(),
                "gemm_bf16: the multi-accumulator weight-gradient kernel needs one work item per CTA");
  }
  p.tail_from = -1;
  p.num_items = p.m_tiles * p.n_tiles * p.splits;
  {
    const char* e = getenv("CSG_GEMM_TAIL_SPLIT");
    const int T = p.m_tiles * p.n_tiles, W = pair ? csg_num_sms() / 2 : csg_num_sms();
    if (!(e && e[0] == '0') && !mn_major && p.smem_epi && p.splits == 1 && N % BN == 0 && BN >= 128 && T > W) {
      const int rem = T % W;
      if (rem > 0 && 2 * rem <= W) { p.tail_from = T - rem; p.num_items = T + rem; }
    }
  }
  int rc;
  if (!mn_major && pair)
    rc = gather == 1 ? (sub32 ? launch_bn<false, G_A, 2, 8, true>(BN, maps, p, stream) : launch_bn<false, G_A, 2>(BN, maps, p, stream))
                     : (sub32 ? launch_bn<false, G_NONE, 2, 8, true>(BN, maps, p, stream) : launch_bn<false, G_NONE, 2>(BN, maps, p, stream));
  else if (!mn_major)
    rc = gather == 1 ? (sub32 ? launch_bn<false, G_A, 1, 8, true>(BN, maps, p, stream) : launch_bn<false, G_A, 1>(BN, maps, p, stream))
                     : (sub32 ? launch_bn<false, G_NONE, 1, 8, true>(BN, maps, p, stream) : launch_bn<false, G_NONE, 1>(BN, maps, p, stream));
  else if (MT == GB_MT) rc = launch<128, true, G_B, 1, GB_MT>(maps, p, stream);
  else rc = gather == 2 ? launch_bn<true, G_B, 1>(BN, maps, p, stream) : launch_bn<true, G_NONE, 1>(BN, maps, p, stream);
  if (rc) return rc;
  if (p.splits > 1) {
    CSG_REQUIRE((long long)M * N < (1ll << 31), "gemm_bf16: split-K output too large");
    job->partial = reinterpret_cast<const float*>(p.C);
    job->out = reinterpret_cast<float*>(out);
    job->n = M * N; job->parts = p.splits; job->stride = (long long)M * N;
  }
  return 0;
}
