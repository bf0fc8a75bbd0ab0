// Geometric ("location") triplets of a scene graph on the device: sg2im/data/base_dataset.py:35-87 of the reference,
// which walks all ordered object pairs in Python inside the DataLoader workers (60-110 ms per graph at 21-26
// objects) and then reduces each of the six augmented relations to its minimal graph (scripts/graphs_utils.py:64-71).
//
// One CTA per graph, one warp per augmented relation (order of BaseDataset.augmented_relations: below, above, left of,
// right of, inside, surrounding).  The warp builds the relation's adjacency bitset from the pairwise box predicates,
// and -- when the relation has at least 3 edges, as triplets_to_minimal demands -- replaces it by the Hsu reduction
// of its Warshall closure (the same row-OR / sequential row-AND-NOT formulation as canon.cu).  Edges are emitted
// relation by relation in row-major (s, o) order, which is the order the reference appends them in.
//
// Float semantics: boxes / centers are float32 and every operation of the reference is a single float32 op
// (x0 + w / 2, c_s - c_o), reproduced here with explicit round-to-nearest intrinsics (no FMA contraction).
//
// The kernel runs twice: COUNT (edges per graph) and EMIT (after csg_canon_offsets' exclusive scan).
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int LOC_REL = 6;
constexpr int LOC_THREADS = LOC_REL * 32;

struct LocParams {
  const float* boxes;      // [NO, 4] xywh
  const float* centers;    // [NO, 2]
  const long long* objs;   // [NO] (stride objs_stride): object class ids
  long long objs_stride;
  const int* obj_off;      // [B + 1]
  long long image_id;      // class id of the __image__ dummy (never takes part)
  int pred[LOC_REL];       // predicate ids of below, above, left of, right of, inside, surrounding
  int W, nmax;
  int* cnt;                // COUNT: [B]
  const int* out_off;      // EMIT: [B + 1]
  long long* out;          // EMIT: [T, 3] graph-local (s, p, o)
};

// bit r of the result: relation r holds for the ordered pair (s, o)   (base_dataset.py:45-79)
__device__ __forceinline__ unsigned pair_relations(float4 bs, float4 bo, float2 cs, float2 co) {
  const float sx1 = __fadd_rn(bs.x, __fdiv_rn(bs.z, 2.f)), sy1 = __fadd_rn(bs.y, __fdiv_rn(bs.w, 2.f));
  const float ox1 = __fadd_rn(bo.x, __fdiv_rn(bo.z, 2.f)), oy1 = __fadd_rn(bo.y, __fdiv_rn(bo.w, 2.f));
  if (bs.x < bo.x && sx1 > ox1 && bs.y < bo.y && sy1 > oy1) return 1u << 5;      // surrounding
  if (bs.x > bo.x && sx1 < ox1 && bs.y > bo.y && sy1 < oy1) return 1u << 4;      // inside
  const float dx = __fsub_rn(cs.x, co.x), dy = __fsub_rn(cs.y, co.y);
  unsigned m = 0u;
  if (dx > 0.f) m |= 1u << 3;          // right of
  else if (dx < 0.f) m |= 1u << 2;     // left of
  if (dy > 0.f) m |= 1u << 0;          // below
  else if (dy < 0.f) m |= 1u << 1;     // above
  return m;
}

template <bool EMIT>
__global__ void __launch_bounds__(LOC_THREADS) location_kernel(LocParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int g = blockIdx.x, rel = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int obeg = p.obj_off[g], n = p.obj_off[g + 1] - obeg;
  const int W = p.W;
  float4* sbox = reinterpret_cast<float4*>(smem_raw);                         // [nmax]
  float2* scen = reinterpret_cast<float2*>(sbox + p.nmax);                    // [nmax]
  unsigned char* real = reinterpret_cast<unsigned char*>(scen + p.nmax);      // [nmax]
  u64* Mall = reinterpret_cast<u64*>(smem_raw + (((size_t)p.nmax * 25 + 15) & ~(size_t)15));   // [6][nmax][W]
  __shared__ int s_cnt[LOC_REL];
  u64* M = Mall + (size_t)rel * p.nmax * W;
  if (n > p.nmax) {                      // sizing error of the caller: flagged through the count, nothing is emitted
    if (!EMIT && threadIdx.x == 0) p.cnt[g] = -1;
    return;
  }
  for (int i = threadIdx.x; i < n; i += LOC_THREADS) {
    sbox[i] = ld_f4(p.boxes + 4 * (size_t)(obeg + i));
    scen[i] = *reinterpret_cast<const float2*>(p.centers + 2 * (size_t)(obeg + i));
    real[i] = (n > 1 && p.objs[(size_t)(obeg + i) * p.objs_stride] != p.image_id) ? 1 : 0;
  }
  for (int i = lane; i < n * W; i += 32) M[i] = 0ull;
  __syncthreads();
  // adjacency of this warp's relation: lane owns rows s = lane, lane + 32, ...
  int count = 0;
  for (int s = lane; s < n; s += 32) {
    if (!real[s]) continue;
    const float4 bs = sbox[s];
    const float2 cs = scen[s];
    for (int o = 0; o < n; ++o) {
      if (o == s || !real[o]) continue;
      if ((pair_relations(bs, sbox[o], cs, scen[o]) >> rel) & 1u) {
        M[s * W + (o >> 6)] |= 1ull << (o & 63);
        ++count;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
  __syncwarp();
  if (count >= 3) {
    // triplets_to_minimal (graphs_utils.py:64-71): Warshall closure in row-OR form, then the sequential Hsu reduction
    for (int i = 0; i < n; ++i) {
      const int iw = i >> 6;
      const u64 ib = 1ull << (i & 63);
      for (int j = lane; j < n; j += 32)
        if (j != i && (M[j * W + iw] & ib))
          for (int w = 0; w < W; ++w) M[j * W + w] |= M[i * W + w];
      __syncwarp();
    }
    for (int j = 0; j < n; ++j) {
      const int jw = j >> 6;
      const u64 jb = 1ull << (j & 63);
      for (int i = lane; i < j; i += 32)
        if (M[i * W + jw] & jb)
          for (int w = 0; w < W; ++w) M[i * W + w] &= ~M[j * W + w];
      __syncwarp();
      if (lane == 0 && (M[j * W + jw] & jb))
        for (int w = 0; w < W; ++w) M[j * W + w] = 0ull;
      __syncwarp();
      for (int i = j + 1 + lane; i < n; i += 32)
        if (M[i * W + jw] & jb)
          for (int w = 0; w < W; ++w) M[i * W + w] &= ~M[j * W + w];
      __syncwarp();
    }
    count = 0;
    for (int i = lane; i < n * W; i += 32) count += __popcll(M[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
  }
  if (lane == 0) s_cnt[rel] = count;
  __syncthreads();
  if (!EMIT) {
    if (threadIdx.x == 0) {
      int t = 0;
      for (int r = 0; r < LOC_REL; ++r) t += s_cnt[r];
      p.cnt[g] = t;
    }
    return;
  }
  int base = p.out_off[g];
  for (int r = 0; r < rel; ++r) base += s_cnt[r];
  // row-major emission: rows in chunks of 32, exclusive scan of the row populations inside the warp
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    int mine = 0;
    if (s < n)
      for (int w = 0; w < W; ++w) mine += __popcll(M[s * W + w]);
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    int pos = base + incl - mine;
    if (s < n) {
      for (int w = 0; w < W; ++w) {
        u64 bits = M[s * W + w];
        while (bits) {
          const int o = (w << 6) + __ffsll((long long)bits) - 1;
          bits &= bits - 1;
          long long* dst = p.out + 3 * (size_t)pos++;
          dst[0] = s; dst[1] = p.pred[rel]; dst[2] = o;
        }
      }
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// add_dummy_triplets (base_dataset.py:141-150): [i, __in_image__, image] for every object i != image, in ascending i,
// appended behind the graph's location triplets.  One warp per graph.  COUNT: cnt[g] = n_g - 1 (0 when the graph holds
// no __image__ object); EMIT: rows at out_off[g] + skip[g].  With several __image__ objects the last one is used
// (the reference's int(nonzero().squeeze()) raises there).
template <bool EMIT>
__global__ void __launch_bounds__(128) dummy_triplets_kernel(const long long* __restrict__ objs, long long objs_stride,
                                                             const int* __restrict__ obj_off, int B, long long image_id,
                                                             int in_image_pred, int* __restrict__ cnt,
                                                             const int* __restrict__ out_off, const int* __restrict__ skip,
                                                             long long* __restrict__ out) {
  CSG_PDL_WAIT();
  const int g = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= B) return;
  const int obeg = obj_off[g], n = obj_off[g + 1] - obeg;
  int img = -1;
  for (int i = lane; i < n; i += 32)
    if (objs[(size_t)(obeg + i) * objs_stride] == image_id) img = i;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) img = max(img, __shfl_xor_sync(0xffffffffu, img, off));
  if (!EMIT) {
    if (lane == 0) cnt[g] = img >= 0 ? n - 1 : 0;
    return;
  }
  if (img < 0) return;
  const size_t base = (size_t)out_off[g] + (skip ? skip[g] : 0);
  for (int i = lane; i < n; i += 32) {
    if (i == img) continue;
    long long* dst = out + 3 * (base + (i < img ? i : i - 1));
    dst[0] = i; dst[1] = in_image_pred; dst[2] = img;
  }
}

int loc_fill(LocParams& p, const float* boxes, const float* centers, const long long* objs, long long objs_stride,
             const int* obj_off, long long image_id, const int* pred_ids, int max_objs, size_t* smem) {
  CSG_REQUIRE(max_objs > 0 && pred_ids, "location_triplets: bad max_objs=%d", max_objs);
  p.boxes = boxes; p.centers = centers; p.objs = objs; p.objs_stride = objs_stride; p.obj_off = obj_off;
  p.image_id = image_id;
  for (int r = 0; r < LOC_REL; ++r) p.pred[r] = pred_ids[r];
  p.W = (max_objs + 63) / 64; p.nmax = max_objs;
  *smem = (((size_t)max_objs * 25 + 15) & ~(size_t)15) + (size_t)LOC_REL * max_objs * p.W * sizeof(u64);
  CSG_REQUIRE(*smem <= 200 * 1024, "location_triplets: %d objects per graph need %zu bytes of shared memory", max_objs, *smem);
  return 0;
}

}  // namespace

// pred_ids: HOST array of the six predicate ids in BaseDataset.augmented_relations order.  cnt[B] receives the number
// of location triplets per graph (a graph with more than max_objs objects gets -1, as csg_canon_count does).
CSG_API int csg_location_count(const float* boxes, const float* centers, const long long* objs, long long objs_stride,
                               const int* obj_off, int B, long long image_id, const int* pred_ids, int max_objs, int* cnt,
                               cudaStream_t stream) {
  if (B == 0) return 0;
  LocParams p;
  size_t smem;
  if (int rc = loc_fill(p, boxes, centers, objs, objs_stride, obj_off, image_id, pred_ids, max_objs, &smem)) return rc;
  p.cnt = cnt; p.out_off = nullptr; p.out = nullptr;
  CSG_CUDA(cudaFuncSetAttribute(location_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  location_kernel<false><<<B, LOC_THREADS, smem, stream>>>(p);
  CSG_CHECK_LAUNCH("csg_location_count");
  return 0;
}

CSG_API int csg_location_emit(const float* boxes, const float* centers, const long long* objs, long long objs_stride,
                              const int* obj_off, int B, long long image_id, const int* pred_ids, int max_objs,
                              const int* out_off, long long* out_triplets, cudaStream_t stream) {
  if (B == 0) return 0;
  LocParams p;
  size_t smem;
  if (int rc = loc_fill(p, boxes, centers, objs, objs_stride, obj_off, image_id, pred_ids, max_objs, &smem)) return rc;
  p.cnt = nullptr; p.out_off = out_off; p.out = out_triplets;
  CSG_CUDA(cudaFuncSetAttribute(location_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  location_kernel<true><<<B, LOC_THREADS, smem, stream>>>(p);
  CSG_CHECK_LAUNCH("csg_location_emit");
  return 0;
}

// add_dummy_triplets (sg2im/data/base_dataset.py:141-150) for a flat batch; see dummy_triplets_kernel above.
CSG_API int csg_dummy_triplets_count(const long long* objs, long long objs_stride, const int* obj_off, int B,
                                     long long image_id, int* cnt, cudaStream_t stream) {
  if (B == 0) return 0;
  CSG_CUDA(csg_launch_pdl(dummy_triplets_kernel<false>, dim3(csg_div_up(B, 4)), dim3(128), 0, stream, objs, objs_stride, obj_off, B,
                          image_id, 0, cnt, (const int*)nullptr, (const int*)nullptr, (long long*)nullptr));
  CSG_CHECK_LAUNCH("csg_dummy_triplets_count");
  return 0;
}

CSG_API int csg_dummy_triplets_emit(const long long* objs, long long objs_stride, const int* obj_off, int B,
                                    long long image_id, int in_image_pred, const int* out_off, const int* skip,
                                    long long* out_triplets, cudaStream_t stream) {
  if (B == 0) return 0;
  CSG_CUDA(csg_launch_pdl(dummy_triplets_kernel<true>, dim3(csg_div_up(B, 4)), dim3(128), 0, stream, objs, objs_stride, obj_off, B,
                          image_id, in_image_pred, (int*)nullptr, out_off, skip, out_triplets));
  CSG_CHECK_LAUNCH("csg_dummy_triplets_emit");
  return 0;
}
