// Index / pooling kernels of the triple graph convolution (sg2im/graph.py:44-113).
//
// All tensors are FLAT: objects [NO, .], triples [NT, .], with per-graph ranges given by
// tri_off[B+1] / obj_off[B+1].  The reference's padded [B, O, .] / [B, T, .] batch is the
// special case tri_off[b] = b*T, obj_off[b] = b*O (padded rows are ordinary rows whose
// `valid` flag is 0: they go through net1 but are not pooled, graph.py:85-107).
//
//   csg_triple_prep      int64 triplets -> int32 global subject/object ids, predicate, type, valid
//   csg_csr_build        two stable CSR orderings of the triples (by subject, by object)
//   csg_triple_conf      confidence = [type==0] + [type==1] * sigmoid(w_trans[pred])   graph.py:69-74
//   csg_segpool_fwd      deterministic confidence-weighted average pooling              graph.py:85-107
//   csg_pool_bwd_obj     dS = dpooled / cnt, dcnt = -<dpooled, pooled> / cnt
//   csg_triple_bwd_assemble   gradient wrt the net1 pre-activation + per-triple dconf
//   csg_segsum           d obj_vecs = segmented sums of the gathered-input gradient     (backward of graph.py:63-64)
//   csg_conf_bwd         d w_trans[p] = sum_t dconf[t] * sigmoid'(w[p])
#include "common.cuh"

namespace {

__device__ __forceinline__ int find_graph(const int* __restrict__ off, int B, int t) {
  int lo = 0, hi = B;   // largest g with off[g] <= t
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// Subject / object ids are graph-local and become raw row offsets of every later kernel (gathers, CSR counters in
// shared memory): an id outside [0, n_g) or a predicate id outside [0, num_preds) -- the inputs at which the
// reference raises IndexError (graph.py:63-64,73,98-103) -- is reported through the asynchronous error record
// and the row is neutralised (ids 0, valid 0, type 2 = zero confidence) so that nothing reads out of bounds.
__device__ __forceinline__ void triple_prep_store(int t, int base, int n_g, long long s, long long p, long long o,
                                                  int ty, int ok_valid, int num_preds, int* __restrict__ s_idx,
                                                  int* __restrict__ o_idx, int* __restrict__ pred, int* __restrict__ type32,
                                                  int* __restrict__ valid, int* err) {
  const bool bad_obj = s < 0 || s >= n_g || o < 0 || o >= n_g;
  const bool bad_pred = num_preds > 0 && (p < 0 || p >= num_preds);
  if (bad_obj || bad_pred) {
    if (bad_obj) csg_report_index(err, CSG_ERR_TRIPLE_OBJECT, t, (s < 0 || s >= n_g) ? s : o, n_g);
    else csg_report_index(err, CSG_ERR_TRIPLE_PREDICATE, t, p, num_preds);
    s_idx[t] = base; o_idx[t] = base; pred[t] = 0; type32[t] = 2; valid[t] = 0;
    return;
  }
  s_idx[t] = base + (int)s;
  o_idx[t] = base + (int)o;
  pred[t] = (int)p;
  type32[t] = ty;
  valid[t] = ok_valid;
}

__global__ void triple_prep_kernel(const long long* __restrict__ triplets, const long long* __restrict__ ttype,
                                   const int* __restrict__ tri_off, const int* __restrict__ obj_off, int B,
                                   int NT, int T_pad, int O_pad, int padding_id, int num_preds,
                                   int* __restrict__ s_idx, int* __restrict__ o_idx, int* __restrict__ pred,
                                   int* __restrict__ type32, int* __restrict__ valid, int* err) {
  CSG_PDL_WAIT();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  int base, n_g;
  if (T_pad > 0) { base = (t / T_pad) * O_pad; n_g = O_pad; }
  else { const int g = find_graph(tri_off, B, t); base = obj_off[g]; n_g = obj_off[g + 1] - base; }
  long long s = triplets[3 * (size_t)t], p = triplets[3 * (size_t)t + 1], o = triplets[3 * (size_t)t + 2];
  triple_prep_store(t, base, n_g, s, p, o, ttype ? (int)ttype[t] : 0, (p != padding_id) ? 1 : 0, num_preds,
                    s_idx, o_idx, pred, type32, valid, err);
}

// GraphTripleConv.forward's own argument layout (graph.py:44): edges [NT, 2], predicate ids, indicators
__global__ void triple_prep_edges_kernel(const long long* __restrict__ edges, const long long* __restrict__ pred_ids,
                                         const unsigned char* __restrict__ indicators,
                                         const long long* __restrict__ ttype, const int* __restrict__ tri_off,
                                         const int* __restrict__ obj_off, int B, int NT, int T_pad, int O_pad,
                                         int num_preds, int* __restrict__ s_idx, int* __restrict__ o_idx,
                                         int* __restrict__ pred, int* __restrict__ type32, int* __restrict__ valid,
                                         int* err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  int base, n_g;
  if (T_pad > 0) { base = (t / T_pad) * O_pad; n_g = O_pad; }
  else { const int g = find_graph(tri_off, B, t); base = obj_off[g]; n_g = obj_off[g + 1] - base; }
  triple_prep_store(t, base, n_g, edges[2 * (size_t)t], pred_ids[t], edges[2 * (size_t)t + 1],
                    ttype ? (int)ttype[t] : 0, indicators ? (indicators[t] ? 1 : 0) : 1, num_preds,
                    s_idx, o_idx, pred, type32, valid, err);
}

__global__ void compose_index_kernel(const int* __restrict__ idx, const long long* __restrict__ map, long long map_stride,
                                     int n, int n_map, int n_values, int* __restrict__ out, int* err) {
  CSG_PDL_WAIT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[i];
  if (j < 0 || j >= n_map) { csg_report_index(err, CSG_ERR_TRIPLE_OBJECT, i, j, n_map); out[i] = 0; return; }
  const long long v = map[(size_t)j * map_stride];
  if (n_values > 0 && (v < 0 || v >= n_values)) { csg_report_index(err, CSG_ERR_EMBED_ID, j, v, n_values); out[i] = 0; return; }
  out[i] = (int)v;
}

__global__ void offsets_uniform_kernel(int* __restrict__ off, int B, int stride) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= B) off[i] = i * stride;
}

// ---------------------------------------------------------------- CSR
__global__ void csr_hist_kernel(const int* __restrict__ ks, const int* __restrict__ ko, int NT,
                                int* __restrict__ cnt_s, int* __restrict__ cnt_o) {
  CSG_PDL_WAIT();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  atomicAdd(&cnt_s[ks[t]], 1);   // integer atomics: the result does not depend on order
  atomicAdd(&cnt_o[ko[t]], 1);
}

// single-CTA exclusive scan of n counts -> row_ptr[n+1]; also copies the starts into cursor[n]
__global__ void __launch_bounds__(1024) csr_scan_kernel(const int* __restrict__ cnt, int n, int* __restrict__ row_ptr,
                                                        int* __restrict__ cursor) {
  CSG_PDL_WAIT();
  __shared__ int sums[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int beg = min(n, tid * per), end = min(n, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += cnt[i];
  sums[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int v = tid >= off ? sums[tid - off] : 0;
    __syncthreads();
    sums[tid] += v;
    __syncthreads();
  }
  int run = tid ? sums[tid - 1] : 0;
  for (int i = beg; i < end; ++i) {
    row_ptr[i] = run;
    cursor[i] = run;
    run += cnt[i];
  }
  if (tid == 1023) row_ptr[n] = sums[1023];
}

// One warp (= one block) per graph walks its triples in order, 32 at a time: lanes with equal keys are ranked
// with match_any, so perm lists every object's triples in ascending triple id (== stable sort).
// The objects of a graph are touched by that graph's warp only; their running cursors live in shared memory
// (falling back to the global cursor array for graphs with more than CSR_SMEM_OBJS objects).
constexpr int CSR_SMEM_OBJS = 2048;
// Blocks [0, B) fill the by-subject ordering, blocks [B, 2B) the by-object ordering (one launch for both).
__global__ void __launch_bounds__(32) csr_fill_kernel(const int* __restrict__ keys_s, const int* __restrict__ keys_o,
                                                      const int* __restrict__ tri_off, const int* __restrict__ obj_off, int B,
                                                      int* __restrict__ cursor_s, int* __restrict__ cursor_o,
                                                      int* __restrict__ perm_s, int* __restrict__ perm_o) {
  CSG_PDL_WAIT();
  __shared__ int scur[CSR_SMEM_OBJS];
  const bool second = blockIdx.x >= B;
  const int g = second ? blockIdx.x - B : blockIdx.x;
  if (g >= B) return;
  const int* __restrict__ keys = second ? keys_o : keys_s;
  int* __restrict__ cursor = second ? cursor_o : cursor_s;
  int* __restrict__ perm = second ? perm_o : perm_s;
  const int lane = threadIdx.x;
  const int beg = tri_off[g], end = tri_off[g + 1];
  const int obeg = obj_off[g], nobj = obj_off[g + 1] - obeg;
  const bool use_smem = nobj <= CSR_SMEM_OBJS;
  if (use_smem) {
    for (int i = lane; i < nobj; i += 32) scur[i] = cursor[obeg + i];
    __syncwarp();
  }
  int key_next = beg + lane < end ? keys[beg + lane] : 0;        // keys are fetched one 32-triple step ahead
  for (int t0 = beg; t0 < end; t0 += 32) {
    int t = t0 + lane;
    bool act = t < end;
    int key = act ? key_next : -1 - lane;
    if (t + 32 < end) key_next = keys[t + 32];
    unsigned peers = __match_any_sync(0xffffffffu, key);
    int rank = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if (act) base = use_smem ? scur[key - obeg] : cursor[key];
    __syncwarp();
    if (act) {
      perm[base + rank] = t;
      if (rank == __popc(peers) - 1) {
        if (use_smem) scur[key - obeg] = base + rank + 1; else cursor[key] = base + rank + 1;
      }
    }
    __syncwarp();
  }
}

__global__ void triple_conf_kernel(const int* __restrict__ type32, const int* __restrict__ pred,
                                   const float* __restrict__ w_trans, int NT, float* __restrict__ conf) {
  CSG_PDL_WAIT();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  int ty = type32[t];
  float c = 0.f;
  if (ty == 0) c = 1.f;
  else if (ty == 1) c = 1.f / (1.f + expf(-w_trans[pred[t]]));
  conf[t] = c;
}

// ---------------------------------------------------------------- pooling
// One CTA of W/4 threads per object; thread j owns columns 4j..4j+3 of the row.
// Order of accumulation = subject incidences in ascending triple id, then object incidences:
// the order CPU scatter_add uses in the reference (graph.py:98-99), so sums are reproducible.
template <bool AVG>
__global__ void segpool_kernel(const float* __restrict__ X, int ldx, int col_s, int col_o, int W,
                               const int* __restrict__ rp_s, const int* __restrict__ perm_s,
                               const int* __restrict__ rp_o, const int* __restrict__ perm_o,
                               const int* __restrict__ valid, const float* __restrict__ conf,
                               float* __restrict__ out, int ldo, float* __restrict__ cnt_out) {
  const int o = blockIdx.x;
  const int c = threadIdx.x * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float cnt = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    const int* rp = pass ? rp_o : rp_s;
    const int* perm = pass ? perm_o : perm_s;
    const int col = (pass ? col_o : col_s) + c;
    const int beg = rp[o], end = rp[o + 1];
    int j = beg;
    for (; j + 4 <= end; j += 4) {   // 4 independent row loads in flight
      int t0 = perm[j], t1 = perm[j + 1], t2 = perm[j + 2], t3 = perm[j + 3];
      bool v0 = !AVG || valid[t0], v1 = !AVG || valid[t1], v2 = !AVG || valid[t2], v3 = !AVG || valid[t3];
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r0 = (v0 && c < W) ? ld_f4(X + (size_t)t0 * ldx + col) : z;
      float4 r1 = (v1 && c < W) ? ld_f4(X + (size_t)t1 * ldx + col) : z;
      float4 r2 = (v2 && c < W) ? ld_f4(X + (size_t)t2 * ldx + col) : z;
      float4 r3 = (v3 && c < W) ? ld_f4(X + (size_t)t3 * ldx + col) : z;
      if (v0) { acc.x += r0.x; acc.y += r0.y; acc.z += r0.z; acc.w += r0.w; if (AVG) cnt += conf[t0]; }
      if (v1) { acc.x += r1.x; acc.y += r1.y; acc.z += r1.z; acc.w += r1.w; if (AVG) cnt += conf[t1]; }
      if (v2) { acc.x += r2.x; acc.y += r2.y; acc.z += r2.z; acc.w += r2.w; if (AVG) cnt += conf[t2]; }
      if (v3) { acc.x += r3.x; acc.y += r3.y; acc.z += r3.z; acc.w += r3.w; if (AVG) cnt += conf[t3]; }
    }
    for (; j < end; ++j) {
      int t = perm[j];
      if (AVG && !valid[t]) continue;
      if (c < W) {
        float4 r = ld_f4(X + (size_t)t * ldx + col);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
      }
      if (AVG) cnt += conf[t];
    }
  }
  if (AVG && cnt > 0.f) {   // graph.py:105-106
    acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt);
    acc.z = __fdiv_rn(acc.z, cnt); acc.w = __fdiv_rn(acc.w, cnt);
  }
  if (c < W) st_f4(out + (size_t)o * ldo + c, acc);
  if (AVG && threadIdx.x == 0) cnt_out[o] = cnt;
}

// dS[o, :] = dpooled[o, :] / cnt[o] (cnt > 0) ; dcnt[o] = -sum_j dpooled[o, j] * pooled[o, j] / cnt[o]
__global__ void pool_bwd_obj_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                                    const float* __restrict__ cnt, int W, float* __restrict__ dS,
                                    float* __restrict__ dcnt) {
  CSG_PDL_WAIT();
  const int o = blockIdx.x;
  const float c = cnt[o];
  const bool nz = c > 0.f;
  float dot = 0.f;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float g = dpooled[(size_t)o * W + j];
    dot += g * pooled[(size_t)o * W + j];
    dS[(size_t)o * W + j] = nz ? __fdiv_rn(g, c) : g;
  }
  __shared__ float red[32];
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += red[w];
    dcnt[o] = nz ? -s / c : 0.f;
  }
}

// One warp per triple.  out = relu(z) * conf was saved by the forward; columns [0,H) subject part,
// [H, H+Dp) predicate part, [H+Dp, 2H+Dp) object part (graph.py:79-81).
//   raw[j]   = d new_t[t, j]  (dS[s_t] | d_newp[t] | dS[o_t]; pooled parts only for valid triples)
//   g[t, j]  = raw[j] * conf[t] * [out[t, j] > 0]                 -> gradient wrt net1's pre-activation
//   dconf[t] = [type==1] * sum_j raw[j] * out[t, j] / conf[t]  +  valid * (dcnt[s_t] + dcnt[o_t])
__global__ void triple_bwd_assemble_kernel(const float* __restrict__ out, const float* __restrict__ dS,
                                           const float* __restrict__ d_newp, const float* __restrict__ dcnt,
                                           const int* __restrict__ s_idx, const int* __restrict__ o_idx,
                                           const int* __restrict__ valid, const int* __restrict__ type32,
                                           const float* __restrict__ conf, int NT, int H, int Dp,
                                           float* __restrict__ g, float* __restrict__ dconf) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= NT) return;
  const int lane = threadIdx.x & 31;
  const int Wd = 2 * H + Dp;
  const int s = s_idx[t], o = o_idx[t];
  const bool v = valid[t] != 0;
  const float cf = conf[t];
  const float* orow = out + (size_t)t * Wd;
  float* grow = g + (size_t)t * Wd;
  float dot = 0.f;
  for (int j = lane * 4; j < Wd; j += 128) {
    float4 raw;
    if (j < H) raw = v ? ld_f4(dS + (size_t)s * H + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    else if (j < H + Dp) raw = d_newp ? ld_f4(d_newp + (size_t)t * Dp + (j - H)) : make_float4(0.f, 0.f, 0.f, 0.f);
    else raw = v ? ld_f4(dS + (size_t)o * H + (j - H - Dp)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y = ld_f4(orow + j);
    dot += raw.x * y.x + raw.y * y.y + raw.z * y.z + raw.w * y.w;
    float4 r;
    r.x = y.x > 0.f ? raw.x * cf : 0.f;
    r.y = y.y > 0.f ? raw.y * cf : 0.f;
    r.z = y.z > 0.f ? raw.z * cf : 0.f;
    r.w = y.w > 0.f ? raw.w * cf : 0.f;
    st_f4(grow + j, r);
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    float dc = 0.f;
    if (type32[t] == 1 && cf > 0.f) dc = dot / cf;
    if (v) dc += dcnt[s] + dcnt[o];
    dconf[t] = dc;
  }
}

// d w_trans[p] += sum over type-1 triples with predicate p of dconf[t] * s(1-s), s = sigmoid(w[p]).
// Deterministic: each warp walks a contiguous chunk in order, lanes with equal predicate are summed
// in lane order by the group leader into warp-private bins; bins are then reduced in fixed order.
constexpr int CONF_BWD_BLOCKS = 256;   // (64 blocks made the partial pass 3x slower: each warp walks its triples serially)
__global__ void conf_bwd_partial_kernel(const float* __restrict__ dconf, const int* __restrict__ type32,
                                        const int* __restrict__ pred, int NT, int P, float* __restrict__ partial) {
  CSG_PDL_WAIT();
  extern __shared__ float bins[];   // [warps][P]
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < warps * P; i += blockDim.x) bins[i] = 0.f;
  __syncthreads();
  const int total_warps = gridDim.x * warps;
  const int gw = blockIdx.x * warps + warp;
  const int per = ((NT + total_warps - 1) / total_warps + 31) / 32 * 32;
  const int beg = min(NT, gw * per), end = min(NT, beg + per);
  float* mybins = bins + warp * P;
  for (int t0 = beg; t0 < end; t0 += 32) {
    int t = t0 + lane;
    bool act = t < end && type32[t] == 1;
    int p = act ? pred[t] : -1 - lane;
    float v = act ? dconf[t] : 0.f;
    unsigned peers = __match_any_sync(0xffffffffu, p);
    bool leader = act && (__ffs(peers) - 1 == lane);
    float sum = 0.f;
    for (int l = 0; l < 32; ++l) {       // fixed lane order
      float x = __shfl_sync(0xffffffffu, v, l);
      if (leader && ((peers >> l) & 1u)) sum += x;
    }
    if (leader) mybins[p] += sum;
    __syncwarp();
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < warps; ++w) s += bins[w * P + p];
    partial[blockIdx.x * P + p] = s;
  }
}
__global__ void conf_bwd_final_kernel(const float* __restrict__ partial, const float* __restrict__ w_trans, int P,
                                      int blocks, float* __restrict__ dw) {
  CSG_PDL_WAIT();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float s = ordered_sum<16>(partial + p, (size_t)P, blocks);
  float sg = 1.f / (1.f + expf(-w_trans[p]));
  dw[p] = s * sg * (1.f - sg);
}

}  // namespace

// ----------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------
CSG_API int csg_compose_index(const int* idx, const long long* map, long long map_stride, int n, int n_map, int n_values,
                              int* out, cudaStream_t stream) {
  if (n == 0) return 0;
  CSG_CUDA(csg_launch_pdl(compose_index_kernel, dim3(csg_div_up(n, 256)), dim3(256), 0, stream, idx, map, map_stride, n, n_map,
                          n_values, out, csg_async_err_ptr()));
  CSG_CHECK_LAUNCH("csg_compose_index");
  return 0;
}

CSG_API int csg_offsets_uniform(int* off, int B, int stride, cudaStream_t stream) {
  offsets_uniform_kernel<<<csg_div_up(B + 1, 256), 256, 0, stream>>>(off, B, stride);
  CSG_CHECK_LAUNCH("csg_offsets_uniform");
  return 0;
}

CSG_API int csg_triple_prep(const long long* triplets, const long long* triplet_type, const int* tri_off,
                            const int* obj_off, int B, int NT, int T_pad, int O_pad, int padding_id, int num_preds,
                            int* s_idx, int* o_idx, int* pred, int* type32, int* valid, cudaStream_t stream) {
  if (NT == 0) return 0;
  CSG_REQUIRE(T_pad > 0 || (tri_off && obj_off), "triple_prep: ragged mode needs offsets");
  CSG_CUDA(csg_launch_pdl(triple_prep_kernel, dim3(csg_div_up(NT, 256)), dim3(256), 0, stream, triplets, triplet_type, tri_off, obj_off, B, NT, T_pad,
                                                              O_pad, padding_id, num_preds, s_idx, o_idx, pred, type32, valid,
                                                              csg_async_err_ptr()));
  CSG_CHECK_LAUNCH("csg_triple_prep");
  return 0;
}

CSG_API int csg_triple_prep_edges(const long long* edges, const long long* pred_ids, const unsigned char* indicators,
                                  const long long* triplet_type, const int* tri_off, const int* obj_off, int B,
                                  int NT, int T_pad, int O_pad, int num_preds, int* s_idx, int* o_idx, int* pred,
                                  int* type32, int* valid, cudaStream_t stream) {
  if (NT == 0) return 0;
  CSG_REQUIRE(T_pad > 0 || (tri_off && obj_off), "triple_prep_edges: ragged mode needs offsets");
  triple_prep_edges_kernel<<<csg_div_up(NT, 256), 256, 0, stream>>>(edges, pred_ids, indicators, triplet_type, tri_off,
                                                                    obj_off, B, NT, T_pad, O_pad, num_preds, s_idx, o_idx,
                                                                    pred, type32, valid, csg_async_err_ptr());
  CSG_CHECK_LAUNCH("csg_triple_prep_edges");
  return 0;
}

CSG_API size_t csg_csr_workspace(int NO) { return (size_t)4 * (NO + 1) * sizeof(int); }

CSG_API int csg_csr_build(const int* keys_s, const int* keys_o, const int* tri_off, const int* obj_off, int B, int NT, int NO,
                          int* rowptr_s, int* perm_s, int* rowptr_o, int* perm_o,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  CSG_REQUIRE(workspace_bytes >= csg_csr_workspace(NO), "csr_build: workspace too small");
  int* cnt_s = reinterpret_cast<int*>(workspace);
  int* cnt_o = cnt_s + (NO + 1);
  int* cur_s = cnt_o + (NO + 1);
  int* cur_o = cur_s + (NO + 1);
  CSG_CUDA(cudaMemsetAsync(cnt_s, 0, (size_t)2 * (NO + 1) * sizeof(int), stream));
  if (NT > 0) {
    CSG_CUDA(csg_launch_pdl(csr_hist_kernel, dim3(csg_div_up(NT, 256)), dim3(256), 0, stream, keys_s, keys_o, NT, cnt_s, cnt_o));
    CSG_CHECK_LAUNCH("csg_csr_build hist");
  }
  CSG_CUDA(csg_launch_pdl(csr_scan_kernel, dim3(1), dim3(1024), 0, stream, cnt_s, NO, rowptr_s, cur_s));
  CSG_CUDA(csg_launch_pdl(csr_scan_kernel, dim3(1), dim3(1024), 0, stream, cnt_o, NO, rowptr_o, cur_o));
  CSG_CHECK_LAUNCH("csg_csr_build scan");
  if (NT > 0 && B > 0) {
    CSG_CUDA(csg_launch_pdl(csr_fill_kernel, dim3(2 * B), dim3(32), 0, stream, keys_s, keys_o, tri_off, obj_off, B, cur_s, cur_o, perm_s, perm_o));
    CSG_CHECK_LAUNCH("csg_csr_build fill");
  }
  return 0;
}

CSG_API int csg_triple_conf(const int* type32, const int* pred, const float* w_trans, int NT, float* conf,
                            cudaStream_t stream) {
  if (NT == 0) return 0;
  CSG_CUDA(csg_launch_pdl(triple_conf_kernel, dim3(csg_div_up(NT, 256)), dim3(256), 0, stream, type32, pred, w_trans, NT, conf));
  CSG_CHECK_LAUNCH("csg_triple_conf");
  return 0;
}

// avg != 0: confidence-weighted average over valid triples (forward pooling); cnt_out[NO] receives the
// denominators.  avg == 0: plain segmented sum over all triples (gather backward).
CSG_API int csg_segpool_f32(const float* X, int ldx, int col_s, int col_o, int W,
                            const int* rowptr_s, const int* perm_s, const int* rowptr_o, const int* perm_o,
                            const int* valid, const float* conf, int NO, float* out, int ldo, float* cnt_out,
                            int avg, cudaStream_t stream) {
  if (NO == 0) return 0;
  CSG_REQUIRE((W & 3) == 0 && (ldx & 3) == 0 && (col_s & 3) == 0 && (col_o & 3) == 0 && (ldo & 3) == 0,
              "segpool: widths/offsets must be multiples of 4");
  CSG_REQUIRE(W <= 4096, "segpool: W=%d too wide", W);
  int threads = ((W / 4 + 31) / 32) * 32;
  if (avg) {
    // valid / conf are read per incidence only: a batch without triples may pass NULL for them
    CSG_REQUIRE(cnt_out, "segpool(avg): cnt_out required");
    segpool_kernel<true><<<NO, threads, 0, stream>>>(X, ldx, col_s, col_o, W, rowptr_s, perm_s, rowptr_o, perm_o,
                                                     valid, conf, out, ldo, cnt_out);
  } else {
    segpool_kernel<false><<<NO, threads, 0, stream>>>(X, ldx, col_s, col_o, W, rowptr_s, perm_s, rowptr_o, perm_o,
                                                      nullptr, nullptr, out, ldo, nullptr);
  }
  CSG_CHECK_LAUNCH("csg_segpool_f32");
  return 0;
}

CSG_API int csg_pool_bwd_obj(const float* dpooled, const float* pooled, const float* cnt, int NO, int W,
                             float* dS, float* dcnt, cudaStream_t stream) {
  if (NO == 0) return 0;
  CSG_CUDA(csg_launch_pdl(pool_bwd_obj_kernel, dim3(NO), dim3(128), 0, stream, dpooled, pooled, cnt, W, dS, dcnt));
  CSG_CHECK_LAUNCH("csg_pool_bwd_obj");
  return 0;
}

CSG_API int csg_triple_bwd_assemble(const float* out, const float* dS, const float* d_newp, const float* dcnt,
                                    const int* s_idx, const int* o_idx, const int* valid, const int* type32,
                                    const float* conf, int NT, int H, int Dp, float* g, float* dconf,
                                    cudaStream_t stream) {
  if (NT == 0) return 0;
  CSG_REQUIRE((H & 3) == 0 && (Dp & 3) == 0, "bwd_assemble: H, Dp must be multiples of 4");
  triple_bwd_assemble_kernel<<<csg_div_up((long long)NT * 32, 256), 256, 0, stream>>>(
      out, dS, d_newp, dcnt, s_idx, o_idx, valid, type32, conf, NT, H, Dp, g, dconf);
  CSG_CHECK_LAUNCH("csg_triple_bwd_assemble");
  return 0;
}

CSG_API size_t csg_conf_bwd_workspace(int P) { return (size_t)CONF_BWD_BLOCKS * P * sizeof(float); }

CSG_API int csg_conf_bwd(const float* dconf, const int* type32, const int* pred, const float* w_trans, int NT, int P,
                         float* dw, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  CSG_REQUIRE(workspace_bytes >= csg_conf_bwd_workspace(P), "conf_bwd: workspace too small");
  CSG_REQUIRE(P > 0 && P <= 4096, "conf_bwd: P=%d out of range", P);
  float* partial = reinterpret_cast<float*>(workspace);
  const int threads = 128;
  CSG_CUDA(csg_launch_pdl(conf_bwd_partial_kernel, dim3(CONF_BWD_BLOCKS), dim3(threads), (threads / 32) * P * sizeof(float), stream, 
      dconf, type32, pred, NT, P, partial));
  CSG_CHECK_LAUNCH("csg_conf_bwd partial");
  CSG_CUDA(csg_launch_pdl(conf_bwd_final_kernel, dim3(csg_div_up(P, 128)), dim3(128), 0, stream, partial, w_trans, P, CONF_BWD_BLOCKS, dw));
  CSG_CHECK_LAUNCH("csg_conf_bwd final");
  return 0;
}
