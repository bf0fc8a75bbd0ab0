// WSGC canonicalization (edge completion) as integer bitset kernels.
//
// Replaces sg2im/data/base_dataset.py:89-139 (add_learnt_triplets) and the helpers in
// scripts/graphs_utils.py:15-155 of the reference, which run as pure-Python O(n^3) loops inside
// the DataLoader workers.  One CTA owns one scene graph; every relation's adjacency matrix is a
// bitset in shared memory (row s = W 64-bit words over the object ids o):
//
//   A[p][s]  unique input edges            (np.unique, base_dataset.py:90 -- duplicates vanish in the bitset)
//   C[r][o] |= bit s  for every sampled converse edge [o, r, s]      (graphs_utils.py:130-155)
//   U = A | C,  closure(U) by Warshall in row-OR form (graphs_utils.py:15-27), T = closure & ~U
//
// Output order is the reference's: type-0 edges = unique rows of (original + converse + meta)
// in lexicographic (s, p, o) order (np.unique, base_dataset.py:130), followed by the type-1
// (transitive) edges relation by relation in row-major order (base_dataset.py:112-122,134-137).
//
// The converse draw of the k-th unique non-meta triple (relations ascending, row-major inside a
// relation -- the order base_dataset.py:98-110 walks them) uses uniforms[uni_off[g] + k]:
// r = vals[rel][#{i : cdf[rel][i] <= u}], which is what legacy numpy RandomState.choice does with
// that double.  The CDF / value tables depend only on the learned weights and are built on the
// host (float64 softmax, graphs_utils.py:132-139).
//
// The kernel runs twice: COUNT (sizes + conv_counts) and EMIT (after an exclusive scan of the
// sizes); recomputing the small bitset work is cheaper than staging it through HBM.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int CTHREADS = 256;

struct CanonParams {
  const long long* triplets;   // [NTin, 3] local (s, p, o)
  const int* tri_off;          // [B+1]
  const int* obj_off;          // [B+1]  (object counts per graph)
  const double* uniforms;      // [>= NTin] draws; graph g starts at tri_off[g]
  const double* cdf;           // [P, ncand]
  const int* vals;             // [P, ncand]
  int ncand;
  int P, meta0, meta1;
  int learned_converse, learned_transitivity;
  int W;                       // 64-bit words per adjacency row
  int nmax;                    // rows reserved per relation in shared memory
  // COUNT outputs
  int* cnt0;                   // [B] number of type-0 edges
  int* cnt1;                   // [B] number of type-1 edges
  int* conv_counts;            // [B, P, P+1]
  // EMIT inputs / outputs
  const int* out_off;          // [B+1]
  long long* out_triplets;     // [NTout, 3]
  long long* out_type;         // [NTout]
  int* err;                    // asynchronous index-error record (common.cuh), may be NULL
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* scratch, int* total) {
  // CTHREADS-wide exclusive scan through shared memory
  const int tid = threadIdx.x;
  scratch[tid] = v;
  __syncthreads();
  for (int off = 1; off < CTHREADS; off <<= 1) {
    int x = tid >= off ? scratch[tid - off] : 0;
    __syncthreads();
    scratch[tid] += x;
    __syncthreads();
  }
  int incl = scratch[tid];
  *total = scratch[CTHREADS - 1];
  __syncthreads();
  return incl - v;
}

template <bool EMIT>
__global__ void __launch_bounds__(CTHREADS) canon_kernel(CanonParams p) {
  CSG_PDL_WAIT();
  extern __shared__ __align__(16) u64 bits[];
  __shared__ int scratch[CTHREADS];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.obj_off[g + 1] - p.obj_off[g];
  const int W = p.W, P = p.P;
  const int rowsz = p.nmax * W;                 // words per relation
  u64* A = bits;
  u64* C = bits + (size_t)P * rowsz;
  const int tbeg = p.tri_off[g], tend = p.tri_off[g + 1];
  auto is_meta = [&](int r) { return r == p.meta0 || r == p.meta1; };
  if (n > p.nmax) {   // caller under-reported max_objs_per_graph: flag instead of overrunning shared memory
    if (!EMIT && tid == 0) { p.cnt0[g] = -1; p.cnt1[g] = 0; }
    return;
  }

  for (int i = tid; i < 2 * P * rowsz; i += CTHREADS) bits[i] = 0ull;
  __syncthreads();
  // ---- unique input edges
  for (int t = tbeg + tid; t < tend; t += CTHREADS) {
    const long long s = p.triplets[3 * (size_t)t], r = p.triplets[3 * (size_t)t + 1], o = p.triplets[3 * (size_t)t + 2];
    if (s >= 0 && s < n && o >= 0 && o < n && r >= 0 && r < P) {
      atomicOr(&A[(size_t)r * rowsz + (int)s * W + ((int)o >> 6)], 1ull << ((int)o & 63));
    } else if (!EMIT) {     // dropped and reported (the reference indexes its adjacency lists with these ids)
      const bool bad_r = r < 0 || r >= P;
      csg_report_index(p.err, CSG_ERR_CANON_TRIPLET, t, bad_r ? r : ((s < 0 || s >= n) ? s : o), bad_r ? P : n);
    }
  }
  __syncthreads();

  // ---- converse sampling
  if (p.learned_converse) {
    // rows (rel, s) in draw order; each thread owns a contiguous chunk of rows
    const int rows = P * n;
    const int per = (rows + CTHREADS - 1) / CTHREADS;
    const int rbeg = min(rows, tid * per), rend = min(rows, rbeg + per);
    int mine = 0;
    for (int e = rbeg; e < rend; ++e) {
      int rel = e / n, s = e % n;
      if (is_meta(rel)) continue;
      for (int w = 0; w < W; ++w) mine += __popcll(A[(size_t)rel * rowsz + s * W + w]);
    }
    int total;
    int k = block_exclusive_scan(mine, scratch, &total);
    for (int e = rbeg; e < rend; ++e) {
      int rel = e / n, s = e % n;
      if (is_meta(rel)) continue;
      for (int w = 0; w < W; ++w) {
        u64 word = A[(size_t)rel * rowsz + s * W + w];
        while (word) {
          int o = w * 64 + __ffsll((long long)word) - 1;
          word &= word - 1;
          double u = p.uniforms[tbeg + k];
          ++k;
          // searchsorted(cdf, u, side='right') clipped to the last candidate: the cumulative sums are non-decreasing,
          // so the count of leading entries <= u is an upper bound by bisection (log2(ncand) dependent loads
          // instead of a linear walk)
          const double* cdf = p.cdf + (size_t)rel * p.ncand;
          int lo = 0, hi = p.ncand - 1;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(cdf + mid) <= u) lo = mid + 1; else hi = mid;
          }
          const int idx = lo;
          int r = p.vals[(size_t)rel * p.ncand + idx];
          if (!EMIT) atomicAdd(&p.conv_counts[((size_t)g * P + rel) * (P + 1) + r], 1);
          if (r != P) atomicOr(&C[(size_t)r * rowsz + o * W + (s >> 6)], 1ull << (s & 63));   // edge [o, r, s]
        }
      }
    }
    __syncthreads();
  }
  // ---- U = A | C for the non-meta relations (kept in A); C becomes the closure workspace
  for (int i = tid; i < P * rowsz; i += CTHREADS) {
    int rel = i / rowsz;
    if (!is_meta(rel)) {
      u64 u = A[i] | C[i];
      A[i] = u;
      C[i] = p.learned_transitivity ? u : 0ull;
    } else {
      C[i] = 0ull;
    }
  }
  __syncthreads();
  // ---- transitive closure, one warp per relation; T = closure & ~U
  if (p.learned_transitivity) {
    for (int rel = warp; rel < P; rel += CTHREADS / 32) {
      if (is_meta(rel)) continue;
      u64* M = C + (size_t)rel * rowsz;
      const u64* U = A + (size_t)rel * rowsz;
      if (n <= 64 && W == 1) {
        // the whole adjacency matrix lives in registers: lane j holds rows j and j + 32 (one 64-bit word each); the
        // pivot row travels by shuffle, so a pivot step is a handful of ALU instructions with no shared-memory round
        // trip and no warp barrier (ncu r02d: the shared-memory form of this loop was 35-40 % of the kernel)
        u64 r0 = lane < n ? M[lane] : 0ull, r1 = lane + 32 < n ? M[lane + 32] : 0ull;
        for (int i = 0; i < n; ++i) {
          const u64 src = i < 32 ? r0 : r1;
          const u64 pivot = __shfl_sync(0xffffffffu, src, i & 31);
          if (lane != i && ((r0 >> i) & 1ull)) r0 |= pivot;
          if (lane + 32 != i && ((r1 >> i) & 1ull)) r1 |= pivot;
        }
        if (lane < n) M[lane] = r0 & ~U[lane];
        if (lane + 32 < n) M[lane + 32] = r1 & ~U[lane + 32];
        __syncwarp();
      } else {
        for (int i = 0; i < n; ++i) {
          const int iw = i >> 6;
          const u64 ib = 1ull << (i & 63);
          for (int j = lane; j < n; j += 32) {
            if (j != i && (M[j * W + iw] & ib)) {
              for (int w = 0; w < W; ++w) M[j * W + w] |= M[i * W + w];
            }
          }
          __syncwarp();
        }
        for (int i = lane; i < n * W; i += 32) M[i] &= ~U[i];
        __syncwarp();
      }
    }
    __syncthreads();
  }

  // ---- type-0 edges: (s, p, o) lexicographic
  const int out_base = EMIT ? p.out_off[g] : 0;
  int n0;
  {
    const int rows = n * P;
    const int per = (rows + CTHREADS - 1) / CTHREADS;
    const int rbeg = min(rows, tid * per), rend = min(rows, rbeg + per);
    int mine = 0;
    for (int e = rbeg; e < rend; ++e) {
      int s = e / P, rel = e % P;
      for (int w = 0; w < W; ++w) mine += __popcll(A[(size_t)rel * rowsz + s * W + w]);
    }
    int pos = block_exclusive_scan(mine, scratch, &n0);
    if (EMIT) {
      for (int e = rbeg; e < rend; ++e) {
        int s = e / P, rel = e % P;
        for (int w = 0; w < W; ++w) {
          u64 word = A[(size_t)rel * rowsz + s * W + w];
          while (word) {
            int o = w * 64 + __ffsll((long long)word) - 1;
            word &= word - 1;
            size_t q = (size_t)out_base + pos++;
            p.out_triplets[3 * q] = s; p.out_triplets[3 * q + 1] = rel; p.out_triplets[3 * q + 2] = o;
            p.out_type[q] = 0;
          }
        }
      }
    }
  }
  // ---- type-1 edges: relation-major, row-major
  int n1;
  {
    const int rows = P * n;
    const int per = (rows + CTHREADS - 1) / CTHREADS;
    const int rbeg = min(rows, tid * per), rend = min(rows, rbeg + per);
    int mine = 0;
    for (int e = rbeg; e < rend; ++e) {
      int rel = e / n, s = e % n;
      for (int w = 0; w < W; ++w) mine += __popcll(C[(size_t)rel * rowsz + s * W + w]);
    }
    int pos = block_exclusive_scan(mine, scratch, &n1);
    if (EMIT) {
      for (int e = rbeg; e < rend; ++e) {
        int rel = e / n, s = e % n;
        for (int w = 0; w < W; ++w) {
          u64 word = C[(size_t)rel * rowsz + s * W + w];
          while (word) {
            int o = w * 64 + __ffsll((long long)word) - 1;
            word &= word - 1;
            size_t q = (size_t)out_base + n0 + pos++;
            p.out_triplets[3 * q] = s; p.out_triplets[3 * q + 1] = rel; p.out_triplets[3 * q + 2] = o;
            p.out_type[q] = 1;
          }
        }
      }
    }
  }
  if (!EMIT && tid == 0) { p.cnt0[g] = n0; p.cnt1[g] = n1; }
}

// ---------------------------------------------------------------- standalone closure / reduction
// adjacency matrices [G, n, n] uint8 -> closure (graphs_utils.py:15-27) and optionally the sequential
// Hsu reduction of the closure (graphs_utils.py:30-38, j outer / i inner, current rows).
__global__ void closure_kernel(const unsigned char* __restrict__ adj, int n, int W, int reduce,
                               unsigned char* __restrict__ out) {
  extern __shared__ __align__(16) u64 M[];   // [n][W]
  const int g = blockIdx.x, lane = threadIdx.x;
  const unsigned char* a = adj + (size_t)g * n * n;
  for (int i = lane; i < n * W; i += 32) M[i] = 0ull;
  __syncwarp();
  for (int r = lane; r < n; r += 32)
    for (int c = 0; c < n; ++c)
      if (a[(size_t)r * n + c]) M[r * W + (c >> 6)] |= 1ull << (c & 63);
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const int iw = i >> 6;
    const u64 ib = 1ull << (i & 63);
    for (int j = lane; j < n; j += 32)
      if (j != i && (M[j * W + iw] & ib))
        for (int w = 0; w < W; ++w) M[j * W + w] |= M[i * W + w];
    __syncwarp();
  }
  if (reduce) {
    for (int j = 0; j < n; ++j) {
      const int jw = j >> 6;
      const u64 jb = 1ull << (j & 63);
      // rows i < j use row j as it is now; then row j itself (cleared if it has a self loop);
      // rows i > j use the possibly cleared row -- exactly the sequential i loop.
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
  }
  unsigned char* o = out + (size_t)g * n * n;
  for (int r = lane; r < n; r += 32)
    for (int c = 0; c < n; ++c) o[(size_t)r * n + c] = (M[r * W + (c >> 6)] >> (c & 63)) & 1ull;
}

int fill(CanonParams& p, const long long* triplets, const int* tri_off, const int* obj_off, const double* uniforms,
         const double* cdf, const int* vals, int ncand, int P, int meta0, int meta1, int learned_converse,
         int learned_transitivity, int max_objs, size_t* smem) {
  CSG_REQUIRE(P > 0 && max_objs > 0, "canon: bad sizes P=%d max_objs=%d", P, max_objs);
  CSG_REQUIRE(!learned_converse || (uniforms && cdf && vals && ncand > 0), "canon: converse tables missing");
  p.triplets = triplets; p.tri_off = tri_off; p.obj_off = obj_off; p.uniforms = uniforms; p.cdf = cdf; p.vals = vals;
  p.ncand = ncand; p.P = P; p.meta0 = meta0; p.meta1 = meta1;
  p.learned_converse = learned_converse; p.learned_transitivity = learned_transitivity;
  p.W = (max_objs + 63) / 64; p.nmax = max_objs;
  p.err = csg_async_err_ptr();
  *smem = (size_t)2 * P * p.nmax * p.W * sizeof(u64);
  CSG_REQUIRE(*smem <= 220 * 1024, "canon: P=%d with %d objects per graph needs %zu bytes of shared memory", P, max_objs, *smem);
  return 0;
}

}  // namespace

CSG_API int csg_canon_count(const long long* triplets, const int* tri_off, const int* obj_off, int B,
                            const double* uniforms, const double* cdf, const int* vals, int ncand,
                            int P, int meta0, int meta1, int learned_converse, int learned_transitivity,
                            int max_objs_per_graph, int* cnt0, int* cnt1, int* conv_counts, cudaStream_t stream) {
  if (B == 0) return 0;
  CanonParams p;
  size_t smem;
  if (int rc = fill(p, triplets, tri_off, obj_off, uniforms, cdf, vals, ncand, P, meta0, meta1, learned_converse,
                    learned_transitivity, max_objs_per_graph, &smem)) return rc;
  p.cnt0 = cnt0; p.cnt1 = cnt1; p.conv_counts = conv_counts;
  p.out_off = nullptr; p.out_triplets = nullptr; p.out_type = nullptr;
  CSG_CUDA(cudaMemsetAsync(conv_counts, 0, (size_t)B * P * (P + 1) * sizeof(int), stream));
  CSG_CUDA(cudaFuncSetAttribute(canon_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CSG_CUDA(csg_launch_pdl(canon_kernel<false>, dim3(B), dim3(CTHREADS), smem, stream, p));
  CSG_CHECK_LAUNCH("csg_canon_count");
  return 0;
}

CSG_API int csg_canon_emit(const long long* triplets, const int* tri_off, const int* obj_off, int B,
                           const double* uniforms, const double* cdf, const int* vals, int ncand,
                           int P, int meta0, int meta1, int learned_converse, int learned_transitivity,
                           int max_objs_per_graph, const int* out_off, long long* out_triplets, long long* out_type,
                           cudaStream_t stream) {
  if (B == 0) return 0;
  CanonParams p;
  size_t smem;
  if (int rc = fill(p, triplets, tri_off, obj_off, uniforms, cdf, vals, ncand, P, meta0, meta1, learned_converse,
                    learned_transitivity, max_objs_per_graph, &smem)) return rc;
  p.cnt0 = nullptr; p.cnt1 = nullptr; p.conv_counts = nullptr;
  p.out_off = out_off; p.out_triplets = out_triplets; p.out_type = out_type;
  CSG_CUDA(cudaFuncSetAttribute(canon_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CSG_CUDA(csg_launch_pdl(canon_kernel<true>, dim3(B), dim3(CTHREADS), smem, stream, p));
  CSG_CHECK_LAUNCH("csg_canon_emit");
  return 0;
}

// out_off[B+1] = exclusive scan of cnt0 + cnt1; summary[0] = total, summary[1] = min(cnt0) (negative = a graph
// exceeded max_objs_per_graph).  One block; the only host read of the canonicalization is `summary`.
namespace {
__global__ void __launch_bounds__(1024) canon_offsets_kernel(const int* __restrict__ cnt0, const int* __restrict__ cnt1,
                                                             int B, int* __restrict__ out_off, int* __restrict__ summary) {
  CSG_PDL_WAIT();
  __shared__ int sums[1024];
  __shared__ int mins[32];
  const int tid = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int beg = min(tid * per, B), end = min(beg + per, B);
  int local = 0, mn = 0x7fffffff;
  for (int i = beg; i < end; ++i) {
    local += max(cnt0[i], 0) + cnt1[i];
    mn = min(mn, cnt0[i]);
  }
  sums[tid] = local;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int v = tid >= off ? sums[tid - off] : 0;
    __syncthreads();
    sums[tid] += v;
    __syncthreads();
  }
  int run = sums[tid] - local;
  for (int i = beg; i < end; ++i) {
    out_off[i] = run;
    run += max(cnt0[i], 0) + cnt1[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) mins[tid >> 5] = mn;
  __syncthreads();
  if (tid == 0) {
    int m = mins[0];
    for (int i = 1; i < 32; ++i) m = min(m, mins[i]);
    out_off[B] = sums[1023];
    summary[0] = sums[1023];
    summary[1] = B > 0 ? m : 0;
  }
}
}  // namespace

CSG_API int csg_canon_offsets(const int* cnt0, const int* cnt1, int B, int* out_off, int* summary, cudaStream_t stream) {
  CSG_CUDA(csg_launch_pdl(canon_offsets_kernel, dim3(1), dim3(1024), 0, stream, cnt0, cnt1, B, out_off, summary));
  CSG_CHECK_LAUNCH("csg_canon_offsets");
  return 0;
}

// closure (reduce = 0) or minimal graph (reduce = 1) of G adjacency matrices [G, n, n] (uint8 0/1).
CSG_API int csg_canon_closure(const unsigned char* adj, int G, int n, int reduce, unsigned char* out,
                              cudaStream_t stream) {
  if (G == 0 || n == 0) return 0;
  int W = (n + 63) / 64;
  size_t smem = (size_t)n * W * sizeof(u64);
  CSG_REQUIRE(smem <= 200 * 1024, "canon_closure: n=%d too large", n);
  CSG_CUDA(cudaFuncSetAttribute(closure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  closure_kernel<<<G, 32, smem, stream>>>(adj, n, W, reduce, out);
  CSG_CHECK_LAUNCH("csg_canon_closure");
  return 0;
}
