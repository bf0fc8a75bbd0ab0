// Native executor of one GraphTripleConv layer (sg2im/graph.py:44-113) on the bf16 tensor-core engine.
//
// The reference runs a layer as ~13*B small ATen launches from Python; the per-stage C ABI of this library
// (csg_gemm_bf16, csg_segpool_bf16, ...) already collapses that to ~30 launches per layer and direction, but
// issuing them one ctypes call at a time leaves the GPU idle more than half of the step.  Here the whole launch
// sequence of a layer's forward (7 launches) or backward (~24) is issued by ONE call from C++, on the caller's
// stream, out of caller-owned memory:
//
//   saved      activations kept for backward + the bf16 copies of the weights (csg_gconv_bf16_saved_bytes)
//   workspace  scratch that is dead when the call returns on the stream (csg_gconv_bf16_workspace)
//
// Dataflow (identical to canonicalsg2im_b200/graph_tc.py, which remains as the staged form used by the tests):
//   forward   cast weights -> conf (graph.py:69-74) -> F1 (gather fused, graph.py:63-67) -> F2 (+bias, ReLU, x conf)
//             -> CSR pooling (graph.py:83-107) -> net2 (graph.py:110)
//   backward  net2 dW/db/dX -> pooling backward -> assemble d(net1 pre-activation) (+ db2, dconf) -> dW2, dhidden ->
//             net1's first Linear from PER-OBJECT SUMS of dhidden (below) -> d w_trans
//
// Backward of the gathered first Linear (graph.py:60-67).  With x_t = [obj[s_t] | pred_t | obj[o_t]] and W1 = [Ws | Wp | Wo]:
//   dW1 = sum_t dh_t^T x_t  =  [ (S dh)^T obj | dh^T pred | (O dh)^T obj ],    db1 = colsum(S dh),
//   d obj = (S dh) Ws + (O dh) Wo,                                              d pred = dh Wp,
// where S dh [NO, H] / O dh [NO, H] are the sums of dhidden over the triples whose subject / object is a given object
// (csg_segsum2_bf16, one pass each over dhidden).  The two T-sized GEMMs left have N = Dp instead of N = 2 Din + Dp and the
// gathered weight-gradient GEMM, the column-sum pass over dhidden and the segmented sums of dX disappear; everything
// per object is a GEMM over NO rows.  Measured on the cfg2 layer (scratch/bwd_alt.py): 144 us against 168 us.  The
// layer that reads the embedding tables directly (n_gather / n_pred > 0) keeps the gathered dataflow (CSG_BWD_SEGSUM=0
// selects it everywhere).
#include "common.cuh"
#include "internal.h"
#include "csg2im.h"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

struct Dims {
  int NT, NO, Din, Dp, H, Dout, Dpo, P;
  int f16;     // 1: forward tensors (inputs, weights, hidden, net1 output, pooled, net2 hidden, output) are fp16; gradients stay bf16
  int n_gather;   // rows of the table the object segments are gathered from (0: `obj` is [NO, Din] and the gather
                  // indices are s_idx / o_idx); > 0: `obj` is an embedding table, index[9] / index[10] hold class ids
  int n_pred;     // > 0: `pred` is a table of n_pred rows gathered by index[11] (predicate ids); 0: [NT, Dp] rows
  int K1() const { return 2 * Din + Dp; }
  int Wd() const { return 2 * H + Dpo; }
};

inline Dims read_dims(const int* d) { return Dims{d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8] ? 1 : 0, d[9], d[10]}; }

inline size_t al(size_t b) { return (b + 255) & ~(size_t)255; }

// net1's first-layer backward from per-object sums of dhidden (see the file comment)
inline bool use_segsum(const Dims& d) {
  if (d.n_gather || d.n_pred) return false;
  const char* e = getenv("CSG_BWD_SEGSUM");
  return !(e && e[0] == '0');
}

// ---- layout of `saved`
struct Saved {
  size_t w1b, w2b, w3b, w4b, w1t, w2t, w3t, w4t, w1so, conf, hidden, out, pooled32, pooled16, cnt, h2, total;
};
Saved plan_saved(const Dims& d, bool need_bwd) {
  Saved s;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  const size_t e1 = (size_t)d.H * d.K1() * 2, e2 = (size_t)d.Wd() * d.H * 2, e3 = (size_t)d.H * d.H * 2,
               e4 = (size_t)d.Dout * d.H * 2;
  s.w1b = take(e1); s.w2b = take(e2); s.w3b = take(e3); s.w4b = take(e4);
  s.w1t = take(need_bwd ? e1 : 0); s.w2t = take(need_bwd ? e2 : 0); s.w3t = take(need_bwd ? e3 : 0);
  s.w4t = take(need_bwd ? e4 : 0);
  s.w1so = take(need_bwd ? (size_t)d.Din * 2 * d.H * 2 : 0);     // [Din, 2H] = [Ws^T | Wo^T]: B operand of d obj = [S dh | O dh] [Ws ; Wo]
  s.conf = take((size_t)(d.NT > 0 ? d.NT : 1) * 4);
  s.hidden = take((size_t)d.NT * d.H * 2);
  s.out = take((size_t)d.NT * d.Wd() * 2);
  s.pooled32 = take((size_t)d.NO * d.H * 4);
  s.pooled16 = take((size_t)d.NO * d.H * 2);
  s.cnt = take((size_t)d.NO * 4);
  s.h2 = take((size_t)d.NO * d.H * 2);
  s.total = o + 256;
  return s;
}

// ---- layout of the backward workspace
// Every producer of partial sums owns its region: the final passes of a layer run together in ONE launch at the end
// of the layer's backward (csg_reduce_multi), so no partial buffer may be reused before that.
struct Work {
  size_t g4, dh2, dpooled, dS, dcnt, g, dhid, dHso, tmp_p, tmp_so, total;
  size_t sk[6], sk_bytes[6];     // split-K partials of dw4, dw3, dw2, dw1 (gathered), dw1 predicate block, dw1 object blocks
  size_t cs[3], cs_bytes[3];     // column-sum partials of db4, db3, db1
  size_t asm_ws, asm_bytes;      // assemble: column sums of g (db2) + per-predicate confidence-gradient bins (d w_trans)
};
Work plan_work(const Dims& d) {
  Work w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
  w.g4 = take((size_t)d.NO * d.Dout * 2);
  w.dh2 = take((size_t)d.NO * d.H * 2);
  w.dpooled = take((size_t)d.NO * d.H * 4);
  w.dS = take((size_t)d.NO * d.H * 4);
  w.dcnt = take((size_t)d.NO * 4);
  w.g = take((size_t)d.NT * d.Wd() * 2);
  w.dhid = take((size_t)d.NT * d.H * 2);
  w.dHso = take((size_t)d.NO * 2 * d.H * 2);          // [S dh | O dh] bf16
  w.tmp_p = take((size_t)d.H * d.Dp * 4);             // dW1 predicate block, contiguous (scattered into dw1 by the final pass)
  w.tmp_so = take((size_t)2 * d.H * d.Din * 4);       // dW1 subject / object blocks
  w.sk_bytes[0] = csg_gemm_bf16_workspace(d.Dout, d.H, d.NO, 1);
  w.sk_bytes[1] = csg_gemm_bf16_workspace(d.H, d.H, d.NO, 1);
  w.sk_bytes[2] = csg_gemm_bf16_workspace(d.Wd(), d.H, d.NT, 1);
  w.sk_bytes[3] = csg_gemm_bf16_workspace(d.H, d.K1(), d.NT, 1);
  w.sk_bytes[4] = csg_gemm_bf16_workspace(d.H, d.Dp, d.NT, 1);
  w.sk_bytes[5] = csg_gemm_bf16_workspace(2 * d.H, d.Din, d.NO, 1);
  for (int i = 0; i < 6; ++i) w.sk[i] = take(w.sk_bytes[i]);
  w.cs_bytes[0] = csg_colsum_bf16_workspace(d.NO, d.Dout);
  w.cs_bytes[1] = csg_colsum_bf16_workspace(d.NO, d.H);
  w.cs_bytes[2] = mx(csg_colsum_bf16_workspace(d.NT, d.H), csg_colsum_bf16_workspace(d.NO, d.H));
  for (int i = 0; i < 3; ++i) w.cs[i] = take(w.cs_bytes[i]);
  w.asm_bytes = csg_triple_bwd_assemble_bf16_deferred_workspace(d.NT, d.H, d.Dpo, d.P);
  w.asm_ws = take(w.asm_bytes);
  w.total = o + 256;
  return w;
}

inline uint8_t* align256(void* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
}

int check_dims(const Dims& d) {
  CSG_REQUIRE(d.NT >= 0 && d.NO >= 0 && d.P > 0, "gconv_bf16: bad sizes NT=%d NO=%d P=%d", d.NT, d.NO, d.P);
  CSG_REQUIRE(d.Din > 0 && d.Dp > 0 && d.H > 0 && d.Dout > 0 && d.Dpo > 0 && d.Din % 64 == 0 && d.Dp % 64 == 0 &&
              d.H % 64 == 0 && d.Dout % 64 == 0 && d.Dpo % 64 == 0,
              "gconv_bf16: feature widths must be positive multiples of 64 (Din=%d Dp=%d H=%d Dout=%d Dpo=%d)",
              d.Din, d.Dp, d.H, d.Dout, d.Dpo);
  return 0;
}

// out = bf16(y > 0 ? dy : 0) with dy fp32 or bf16 (ReLU backward of the layer output, graph.py:110)
template <bool DY_BF16>
__global__ void relu_mask_out_kernel(const void* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ out, long long n) {
  CSG_PDL_WAIT();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = DY_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy)[i])
                          : reinterpret_cast<const float*>(dy)[i];
  out[i] = __float2bfloat16_rn(pos16(reinterpret_cast<const unsigned short*>(y)[i]) ? v : 0.f);
}

#define CSG_TRY(call)            \
  do {                           \
    if (int rc__ = (call)) return rc__; \
  } while (0)

}  // namespace

// 16-bit copies of a layer's weights (+ the transposes / column blocks backward consumes) laid out as the head of `saved`:
// appends the jobs of one layer to the arrays of a csg_cast_bf16_multi_ld call
static int add_cast_jobs(const Dims& d, const void* const* params, uint8_t* wv, bool need_bwd, int n, const void** src, void** dst,
                  int* rows, int* cols, int* tr, int* lds, int* ldd) {
  const Saved s = plan_saved(d, true);
  const float* w[4] = {(const float*)params[0], (const float*)params[2], (const float*)params[4], (const float*)params[6]};
  const int K1 = d.K1(), Wd = d.Wd();
  __nv_bfloat16* w1so = reinterpret_cast<__nv_bfloat16*>(wv + s.w1so);
  // jobs 8, 9: [Ws^T | Wo^T] = the transposes of the subject / object column blocks of w1, side by side
  const void* jsrc[10] = {w[0], w[1], w[2], w[3], w[0], w[1], w[2], w[3], w[0], w[0] + d.Din + d.Dp};
  void* jdst[10] = {wv + s.w1b, wv + s.w2b, wv + s.w3b, wv + s.w4b, wv + s.w1t, wv + s.w2t, wv + s.w3t, wv + s.w4t, w1so, w1so + d.H};
  const int jrows[10] = {d.H, Wd, d.H, d.Dout, d.H, Wd, d.H, d.Dout, d.H, d.H};
  const int jcols[10] = {K1, d.H, d.H, d.H, K1, d.H, d.H, d.H, d.Din, d.Din};
  const int jtr[10] = {0, 0, 0, 0, 1, 1, 1, 1, 1, 1};
  const int jlds[10] = {0, 0, 0, 0, 0, 0, 0, 0, K1, K1};
  const int jldd[10] = {0, 0, 0, 0, 0, 0, 0, 0, 2 * d.H, 2 * d.H};
  const int count = need_bwd ? (use_segsum(d) ? 10 : 8) : 4;
  for (int i = 0; i < count; ++i, ++n) {
    src[n] = jsrc[i]; dst[n] = jdst[i]; rows[n] = jrows[i]; cols[n] = jcols[i]; tr[n] = jtr[i]; lds[n] = jlds[i]; ldd[n] = jldd[i];
  }
  return n;
}

CSG_API size_t csg_gconv_bf16_saved_bytes(const int* dims, int need_bwd) {
  return plan_saved(read_dims(dims), need_bwd != 0).total;
}
CSG_API size_t csg_gconv_bf16_out_offset(const int* dims, int need_bwd) {
  return plan_saved(read_dims(dims), need_bwd != 0).out;
}
CSG_API size_t csg_gconv_bf16_workspace(const int* dims) { return plan_work(read_dims(dims)).total; }
// bytes of a caller-kept buffer of weight copies (the head of the `saved` layout)
CSG_API size_t csg_gconv_bf16_wbuf_bytes(const int* dims) { return plan_saved(read_dims(dims), true).conf + 256; }
// The 16-bit weight copies of n layers (dims: n x 11 ints, params: n x 9 device pointers as in csg_gconv_bf16_fwd, wbufs: n
// buffers of csg_gconv_bf16_wbuf_bytes, 256-byte aligned) in ONE launch: once per optimizer step instead of once per
// layer and forward.
CSG_API int csg_gconv_bf16_cast_weights(int n, const int* dims, const void* const* params, void* const* wbufs, int need_bwd,
                                        csg_stream_t stream) {
  CSG_REQUIRE(n >= 0 && n <= 6, "gconv_bf16_cast_weights: %d layers (at most 6 per call)", n);
  const void* src[60]; void* dst[60]; int rows[60], cols[60], tr[60], lds[60], ldd[60];
  int jobs = 0, f16 = 0;
  for (int l = 0; l < n; ++l) {
    const Dims d = read_dims(dims + 11 * l);
    CSG_TRY(check_dims(d));
    CSG_REQUIRE(wbufs[l] && (reinterpret_cast<uintptr_t>(wbufs[l]) & 255) == 0, "gconv_bf16_cast_weights: wbuf %d must be 256-byte aligned", l);
    CSG_REQUIRE(l == 0 || d.f16 == f16, "gconv_bf16_cast_weights: mixed formats");
    f16 = d.f16;
    jobs = add_cast_jobs(d, params + 9 * l, reinterpret_cast<uint8_t*>(wbufs[l]), need_bwd != 0, jobs, src, dst, rows, cols, tr, lds, ldd);
  }
  return csg_cast_bf16_multi_ld(jobs, src, dst, rows, cols, tr, lds, ldd, f16, reinterpret_cast<cudaStream_t>(stream));
}
// columns of the dX matrix csg_gconv_bf16_bwd writes: Dp (d pred only) when net1's first Linear is differentiated through the
// per-object sums of dhidden, 2 Din + Dp (the whole gathered row, d pred in columns Din .. Din+Dp) on the gathered dataflow
CSG_API int csg_gconv_bf16_dx_cols(const int* dims) {
  const Dims d = read_dims(dims);
  return use_segsum(d) ? d.Dp : d.K1();
}

// dims (HOST int[11]): {NT, NO, Din, Dp, H, Dout, Dpo, P, fwd_fp16, n_gather, n_pred}; index (HOST void*[12]).  params (HOST array of 9 device pointers, fp32): w1 [H, 2Din+Dp],
// b1, w2 [2H+Dpo, H], b2, w3 [H, H], b3, w4 [Dout, H], b4, w_trans [P].  index (HOST array of 9 device pointers,
// int32): s_idx, o_idx, pred_id, type32, valid [NT]; rowptr_s [NO+1], perm_s [NT], rowptr_o, perm_o.
// obj [NO, Din] bf16 contiguous; pred [NT, Dp] bf16 with row pitch ldp; new_obj [NO, Dout] bf16 (written);
// `saved` must be 256-byte aligned; new_p = saved + csg_gconv_bf16_out_offset, rows of pitch 2H+Dpo, columns H..H+Dpo.
CSG_API int csg_gconv_bf16_fwd(const int* dims, const void* obj, const void* pred, int ldp,
                               const void* const* params, const void* const* index, int need_bwd, void* saved,
                               size_t saved_bytes, void* new_obj, const void* wbuf, const float* conf_ext,
                               csg_stream_t stream) {
  const Dims d = read_dims(dims);
  CSG_TRY(check_dims(d));
  CSG_REQUIRE(!(d.f16 && need_bwd), "gconv_bf16_fwd: fp16 forward tensors (dims[8] = 1) are inference-only");
  const Saved s = plan_saved(d, need_bwd != 0);
  CSG_REQUIRE(saved && saved_bytes >= s.total && (reinterpret_cast<uintptr_t>(saved) & 255) == 0,
              "gconv_bf16_fwd: `saved` must be 256-byte aligned and hold %zu bytes", s.total);
  uint8_t* sv = reinterpret_cast<uint8_t*>(saved);
  const float* w[4] = {(const float*)params[0], (const float*)params[2], (const float*)params[4], (const float*)params[6]};
  const float* b[4] = {(const float*)params[1], (const float*)params[3], (const float*)params[5], (const float*)params[7]};
  const float* w_trans = (const float*)params[8];
  const int* s_idx = (const int*)index[0];
  const int* o_idx = (const int*)index[1];
  const int* pred_id = (const int*)index[2];
  const int* type32 = (const int*)index[3];
  const int* valid = (const int*)index[4];
  const int *rowptr_s = (const int*)index[5], *perm_s = (const int*)index[6], *rowptr_o = (const int*)index[7],
            *perm_o = (const int*)index[8];
  const int K1 = d.K1(), Wd = d.Wd();
  const int fmt = d.f16 ? 7 : 0;      // csg_gemm_bf16 formats: A, B and C of every forward GEMM are forward tensors
  // ---- 16-bit copies of the weights (+ transposes for the dX-type GEMMs of backward), one launch
  // wbuf != NULL: the caller keeps the weight copies (csg_gconv_bf16_cast_weights, once per optimizer step for all
  // layers); conf_ext != NULL: the triple confidences were computed once for all the layers that share w_trans
  const uint8_t* wv = wbuf ? reinterpret_cast<const uint8_t*>(wbuf) : sv;
  if (!wbuf) {
    const void* src[10]; void* dst[10]; int rows[10], cols[10], tr[10], lds[10], ldd[10];
    const int n = add_cast_jobs(d, params, sv, need_bwd != 0, 0, src, dst, rows, cols, tr, lds, ldd);
    CSG_TRY(csg_cast_bf16_multi_ld(n, src, dst, rows, cols, tr, lds, ldd, d.f16, reinterpret_cast<cudaStream_t>(stream)));
  }
  const float* conf = conf_ext;
  if (!conf) {
    float* own = reinterpret_cast<float*>(sv + s.conf);
    CSG_TRY(csg_triple_conf(type32, pred_id, w_trans, d.NT, own, stream));
    conf = own;
  }
  // ---- net1 on the gathered triple rows
  CSG_TRY(csg_gemm_bf16(0, 1, d.NT, d.H, K1, nullptr, 0, wv + s.w1b, K1, sv + s.hidden, d.H, 0, b[0], 1, nullptr, nullptr, 0,
                        obj, pred, d.n_gather ? (const int*)index[9] : s_idx, d.n_gather ? (const int*)index[10] : o_idx,
                        d.Din, d.Dp, ldp, d.n_gather ? d.n_gather : d.NO, d.n_pred ? (const int*)index[11] : nullptr,
                        d.n_pred, fmt, nullptr, 0, stream));
  CSG_TRY(csg_gemm_bf16(0, 0, d.NT, Wd, d.H, sv + s.hidden, d.H, wv + s.w2b, d.H, sv + s.out, Wd, 0, b[1], 1, conf, nullptr, 0,
                        nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, fmt, nullptr, 0, stream));
  // ---- confidence-weighted average onto objects
  CSG_TRY(csg_segpool_bf16(sv + s.out, Wd, 0, d.H + d.Dpo, d.H, rowptr_s, perm_s, rowptr_o, perm_o, valid, conf, d.NO,
                           reinterpret_cast<float*>(sv + s.pooled32), sv + s.pooled16, d.H,
                           reinterpret_cast<float*>(sv + s.cnt), 1, d.f16, stream));
  // ---- net2
  CSG_TRY(csg_gemm_bf16(0, 0, d.NO, d.H, d.H, sv + s.pooled16, d.H, wv + s.w3b, d.H, sv + s.h2, d.H, 0, b[2], 1, nullptr,
                        nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, fmt, nullptr, 0, stream));
  CSG_TRY(csg_gemm_bf16(0, 0, d.NO, d.Dout, d.H, sv + s.h2, d.H, wv + s.w4b, d.H, new_obj, d.Dout, 0, b[3], 1, nullptr,
                        nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, fmt, nullptr, 0, stream));
  return 0;
}

// Backward of the call above.  d_new_obj [NO, Dout] (fp32, or bf16 when d_new_obj_bf16; NULL = zero),
// d_new_p [NT, Dpo] bf16 with row pitch ld_dnewp (NULL = zero).  Written: dobj [NO, Din] (fp32, or bf16 when
// dobj_bf16), dX [NT, csg_gconv_bf16_dx_cols] bf16 (d pred, or the whole gathered row with d pred in its columns
// Din..Din+Dp), dparams (fp32, contiguous, in this order:
// dw1 [H, 2Din+Dp], db1 [H], dw2 [2H+Dpo, H], db2, dw3 [H, H], db3, dw4 [Dout, H], db4, dw_trans [P]).
CSG_API int csg_gconv_bf16_bwd(const int* dims, const void* obj, const void* pred, int ldp,
                               const void* const* params, const void* const* index,
                               const void* d_new_obj, int d_new_obj_bf16, const void* d_new_p, int ld_dnewp,
                               const void* saved, const void* new_obj, void* dobj, int dobj_bf16, void* dX,
                               float* dparams, void* workspace, size_t workspace_bytes, const void* wbuf,
                               const float* conf_ext, csg_stream_t stream) {
  const Dims d = read_dims(dims);
  CSG_TRY(check_dims(d));
  // tcgen05.mma kind::f16 takes ONE 16-bit format for both operands (a mixed fp16 x bf16 instruction descriptor raises
  // "illegal instruction" on sm_100a: scratch/probe_mixed_mma.py), gradients need bf16's range, and every backward GEMM
  // multiplies a gradient with a forward tensor: fp16 forward tensors are therefore inference-only
  CSG_REQUIRE(!d.f16, "gconv_bf16_bwd: fp16 forward tensors (dims[8] = 1) are inference-only");
  const Saved s = plan_saved(d, true);
  const Work w = plan_work(d);
  CSG_REQUIRE(saved && (reinterpret_cast<uintptr_t>(saved) & 255) == 0, "gconv_bf16_bwd: `saved` must be 256-byte aligned");
  CSG_REQUIRE(workspace && workspace_bytes >= w.total, "gconv_bf16_bwd: workspace of %zu bytes needed", w.total);
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  uint8_t* ws = align256(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* w_trans = (const float*)params[8];
  const int* s_idx = (const int*)index[0];
  const int* o_idx = (const int*)index[1];
  const int* pred_id = (const int*)index[2];
  const int* type32 = (const int*)index[3];
  const int* valid = (const int*)index[4];
  const int *rowptr_s = (const int*)index[5], *perm_s = (const int*)index[6], *rowptr_o = (const int*)index[7],
            *perm_o = (const int*)index[8];
  const int K1 = d.K1(), Wd = d.Wd(), H = d.H, NT = d.NT, NO = d.NO, Dout = d.Dout;
  // flat parameter gradients
  float* dw1 = dparams;
  float* db1 = dw1 + (size_t)H * K1;
  float* dw2 = db1 + H;
  float* db2 = dw2 + (size_t)Wd * H;
  float* dw3 = db2 + Wd;
  float* db3 = dw3 + (size_t)H * H;
  float* dw4 = db3 + H;
  float* db4 = dw4 + (size_t)Dout * H;
  float* dwt = db4 + Dout;
  CSG_REQUIRE((reinterpret_cast<uintptr_t>(dparams) & 15) == 0, "gconv_bf16_bwd: dparams must be 16-byte aligned");

  void* g4 = ws + w.g4;
  void* dh2 = ws + w.dh2;
  float* dpooled = reinterpret_cast<float*>(ws + w.dpooled);
  float* dS = reinterpret_cast<float*>(ws + w.dS);
  float* dcnt = reinterpret_cast<float*>(ws + w.dcnt);
  void* g = ws + w.g;
  void* dhid = ws + w.dhid;
  const void* hidden = sv + s.hidden;
  const void* out = sv + s.out;
  const void* h2 = sv + s.h2;
  const void* pooled16 = sv + s.pooled16;
  const float* conf = conf_ext ? conf_ext : reinterpret_cast<const float*>(sv + s.conf);
  const uint8_t* wv = wbuf ? reinterpret_cast<const uint8_t*>(wbuf) : sv;     // weight copies: the caller's, or the head of `saved`
  CsgReduceJob jobs[CSG_REDUCE_MAX_JOBS];
  int njobs = 0;
  // gather indices of the triple input (the embedding tables' class / predicate ids when layer 0 reads them directly)
  const int* g_s = d.n_gather ? (const int*)index[9] : s_idx;
  const int* g_o = d.n_gather ? (const int*)index[10] : o_idx;
  const int* g_p = d.n_pred ? (const int*)index[11] : nullptr;

  // every backward GEMM multiplies a gradient (A, bf16) with a forward tensor (B: activations, weights or the gathered
  // triple input) and writes a gradient (bf16 / fp32); split-K final passes are deferred to the end of the layer
  const int bfmt = 0;
#define GEMM(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, mask, ldm, SK)                                         \
  CSG_TRY(csg_gemm_bf16_deferred(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, nullptr, 0, nullptr, mask, ldm,     \
                                 (gather) ? obj : nullptr, (gather) ? pred : nullptr, (gather) ? g_s : nullptr,        \
                                 (gather) ? g_o : nullptr, (gather) ? d.Din : 0, (gather) ? d.Dp : 0,                  \
                                 (gather) ? ldp : 0, (gather) ? (d.n_gather ? d.n_gather : NO) : 0,                    \
                                 (gather) ? g_p : nullptr, (gather) ? d.n_pred : 0, bfmt, (SK) >= 0 ? ws + w.sk[(SK) >= 0 ? (SK) : 0] : nullptr, \
                                 (SK) >= 0 ? w.sk_bytes[(SK) >= 0 ? (SK) : 0] : 0, st, &jobs[njobs]));             \
  if (jobs[njobs].parts > 0) ++njobs

  // ---- net2 backward (graph.py:110)
  const long long n4 = (long long)NO * Dout;
  if (n4 > 0) {
    if (!d_new_obj) {
      CSG_CUDA(cudaMemsetAsync(g4, 0, (size_t)n4 * 2, st));
    } else if (d_new_obj_bf16) {
      CSG_CUDA(csg_launch_pdl(relu_mask_out_kernel<true>, dim3(csg_div_up(n4, 256)), dim3(256), 0, st, d_new_obj, (const __nv_bfloat16*)new_obj,
                                                                      (__nv_bfloat16*)g4, n4));
      CSG_CHECK_LAUNCH("csg_gconv_bf16_bwd relu mask");
    } else {
      CSG_CUDA(csg_launch_pdl(relu_mask_out_kernel<false>, dim3(csg_div_up(n4, 256)), dim3(256), 0, st, d_new_obj, (const __nv_bfloat16*)new_obj,
                                                                       (__nv_bfloat16*)g4, n4));
      CSG_CHECK_LAUNCH("csg_gconv_bf16_bwd relu mask");
    }
  }
  // (the bias gradients db4 = colsum(g4), db3 = colsum(dh2), db1 are taken in ONE launch further down)
  GEMM(1, 0, Dout, H, NO, g4, Dout, h2, H, dw4, H, 1, nullptr, 0, 0);
  GEMM(0, 0, NO, H, Dout, g4, Dout, wv + s.w4t, Dout, dh2, H, 0, h2, H, -1);
  GEMM(1, 0, H, H, NO, dh2, H, pooled16, H, dw3, H, 1, nullptr, 0, 1);
  GEMM(0, 0, NO, H, H, dh2, H, wv + s.w3t, H, dpooled, H, 1, nullptr, 0, -1);
  // ---- pooling backward (graph.py:83-107)
  CSG_TRY(csg_pool_bwd_obj(dpooled, reinterpret_cast<const float*>(sv + s.pooled32),
                           reinterpret_cast<const float*>(sv + s.cnt), NO, H, dS, dcnt, stream));
  // gradient wrt net1's pre-activation (+ partial column sums = db2, + per-predicate bins of d conf = d w_trans)
  CSG_TRY(csg_triple_bwd_assemble_bf16_deferred(out, dS, d_new_p, d_new_p ? ld_dnewp : 0, dcnt, s_idx, o_idx, valid, type32,
                                                pred_id, conf, w_trans, NT, H, d.Dpo, d.P, g, db2, dwt, d.f16,
                                                ws + w.asm_ws, w.asm_bytes, st, &jobs[njobs], &jobs[njobs + 1]));
  if (jobs[njobs].parts > 0) { if (jobs[njobs + 1].parts > 0) { njobs += 2; } else { ++njobs; } }
  // ---- net1 backward (graph.py:63-67)
  GEMM(1, 0, Wd, H, NT, g, Wd, hidden, H, dw2, H, 1, nullptr, 0, 2);
  GEMM(0, 0, NT, H, Wd, g, Wd, wv + s.w2t, Wd, dhid, H, 0, hidden, H, -1);
  // db4, db3 and db1 (column sums of g4, dh2 and of dhidden or its per-subject sums) in one launch
  auto colsums = [&](const void* X1, int M1, int ld1) -> int {
    const void* X[3] = {g4, dh2, X1};
    const int Ms[3] = {NO, NO, M1}, Ns[3] = {Dout, H, H}, lds[3] = {Dout, H, ld1};
    float* outs[3] = {db4, db3, db1};
    void* wsp[3] = {ws + w.cs[0], ws + w.cs[1], ws + w.cs[2]};
    const size_t wsb[3] = {w.cs_bytes[0], w.cs_bytes[1], w.cs_bytes[2]};
    CsgReduceJob cj[3];
    if (int rc = csg_colsum_bf16_multi_deferred(3, X, Ms, Ns, lds, outs, wsp, wsb, st, cj)) return rc;
    for (int i = 0; i < 3; ++i)
      if (cj[i].parts > 0) jobs[njobs++] = cj[i];
    return 0;
  };
  if (!use_segsum(d)) {
    GEMM(1, 2, H, K1, NT, dhid, H, nullptr, 0, dw1, K1, 1, nullptr, 0, 3);
    CSG_TRY(colsums(dhid, NT, H));
    GEMM(0, 0, NT, K1, H, dhid, H, wv + s.w1t, H, dX, K1, 0, nullptr, 0, -1);
    // ---- gather backward: segmented sums of dX over ALL triples onto their subject / object rows
    CSG_TRY(csg_segpool_bf16(dX, K1, 0, d.Din + d.Dp, d.Din, rowptr_s, perm_s, rowptr_o, perm_o, nullptr, nullptr, NO,
                             dobj_bf16 ? nullptr : reinterpret_cast<float*>(dobj), dobj_bf16 ? dobj : nullptr, d.Din, nullptr,
                             0, 0, stream));
  } else {
    // ---- per-object sums of dhidden: [S dh | O dh] (bf16, [NO, 2H])
    void* dHso = ws + w.dHso;
    float* tmp_p = reinterpret_cast<float*>(ws + w.tmp_p);
    float* tmp_so = reinterpret_cast<float*>(ws + w.tmp_so);
    const int Din = d.Din, Dp = d.Dp;
    CSG_TRY(csg_segsum2_bf16(dhid, H, H, rowptr_s, perm_s, rowptr_o, perm_o, NO, nullptr, dHso, 2 * H, stream));
    // db1 = colsum(dhidden) = colsum(S dh): every triple has exactly one subject
    CSG_TRY(colsums(dHso, NO, 2 * H));
    // a weight-gradient block computed into a contiguous scratch matrix and placed into dw1 (row pitch K1) by the final pass
    auto place = [&](CsgReduceJob& j, const float* scratch, int rows, int cols, float* dst) {
      if (j.parts == 0) {                         // written directly (no split-K): the final pass is then a copy
        j.partial = scratch; j.parts = 1; j.stride = (long long)rows * cols; j.lanes = 1; j.op = CSG_RED_SUM; j.aux = nullptr;
      }
      j.n = rows * cols; j.out = dst; j.ncols = cols; j.ldo = K1;
    };
    // dW1[:, Din : Din+Dp] = dh^T pred
    CSG_TRY(csg_gemm_bf16_deferred(1, 0, H, Dp, NT, dhid, H, pred, ldp, tmp_p, Dp, 1, nullptr, 0, nullptr, nullptr, 0, nullptr,
                                   nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, bfmt, ws + w.sk[4], w.sk_bytes[4], st,
                                   &jobs[njobs]));
    place(jobs[njobs], tmp_p, H, Dp, dw1 + Din);
    ++njobs;
    // dW1[:, 0 : Din] = (S dh)^T obj and dW1[:, Din+Dp :] = (O dh)^T obj: one GEMM with M = 2H, two placements
    CSG_TRY(csg_gemm_bf16_deferred(1, 0, 2 * H, Din, NO, dHso, 2 * H, obj, Din, tmp_so, Din, 1, nullptr, 0, nullptr, nullptr, 0,
                                   nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, bfmt, ws + w.sk[5], w.sk_bytes[5], st,
                                   &jobs[njobs]));
    {
      CsgReduceJob lo = jobs[njobs], hi = jobs[njobs];
      const float* base_partial = lo.parts > 0 ? lo.partial : tmp_so;
      place(lo, tmp_so, H, Din, dw1);
      place(hi, tmp_so, H, Din, dw1 + Din + Dp);
      lo.partial = base_partial;
      hi.partial = base_partial + (size_t)H * Din;
      jobs[njobs] = lo; jobs[njobs + 1] = hi;
      njobs += 2;
    }
    // d pred = dh Wp  (rows Din .. Din+Dp of the transposed copy of w1)
    CSG_TRY(csg_gemm_bf16_deferred(0, 0, NT, Dp, H, dhid, H, wv + s.w1t + (size_t)Din * H * 2, H, dX, Dp, 0, nullptr, 0, nullptr,
                                   nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, bfmt, nullptr, 0, st,
                                   &jobs[njobs]));
    // d obj = [S dh | O dh] [Ws ; Wo]
    CSG_TRY(csg_gemm_bf16_deferred(0, 0, NO, Din, 2 * H, dHso, 2 * H, wv + s.w1so, 2 * H, dobj, Din, dobj_bf16 ? 0 : 1, nullptr, 0,
                                   nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, bfmt, nullptr, 0,
                                   st, &jobs[njobs]));
  }
#undef GEMM
  // ---- all deferred final passes of the layer: dw1..dw4 (split-K), db1..db4 (column sums), d w_trans
  CSG_TRY(csg_reduce_multi(jobs, njobs, st));
  return 0;
}
