// Library-internal interfaces (not part of the C ABI): deferred ordered reductions.
//
// Several backward kernels of a GraphTripleConv layer end in "sum P partial rows in a fixed order": the split-K weight
// gradient GEMMs, the column sums that give the bias gradients, the per-predicate sums of the confidence gradient.
// Run one by one that is ~10 tiny launches per layer whose cost is launch latency.  The *_deferred variants below write
// their partials and describe the missing final pass as a CsgReduceJob; csg_reduce_multi then finishes up to
// CSG_REDUCE_MAX_JOBS of them in ONE launch, in exactly the order the stand-alone entry points use (which are the
// same kernels with a single job), so results are bit-identical either way.
#pragma once
#include "common.cuh"

struct CsgReduceJob {
  const float* partial;   // [parts][stride] fp32, the first n entries of every row are summed
  float* out;             // [n]
  int n, parts;
  long long stride;
  int lanes;              // 1: one thread per output sums the parts in order (8 loads in flight);
                          // 8: eight threads per output take parts y, y + 8, ... in order and are combined in lane order
  int op;                 // CSG_RED_SUM, or CSG_RED_SIGMOID_GRAD: out = sum * s (1 - s), s = sigmoid(aux[i])
  const float* aux;
  int ncols, ldo;         // ncols > 0: the n results form rows of ncols that are written with row pitch ldo
                          // (out[(i / ncols) * ldo + i % ncols]): a column block of a wider matrix; 0: out[i]
};
enum { CSG_RED_SUM = 0, CSG_RED_SIGMOID_GRAD = 1 };
constexpr int CSG_REDUCE_MAX_JOBS = 16;

int csg_reduce_multi(const CsgReduceJob* jobs, int njobs, cudaStream_t stream);

// csg_gemm_bf16 whose split-K final pass is left to the caller: *job describes it (job->parts == 0: nothing to do,
// the result was written directly).  The split-K workspace must stay untouched until the job has run.
int csg_gemm_bf16_deferred(int mn_major, int gather, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                           void* C, int ldc, int out_f32, const float* bias, int relu, const float* rowscale,
                           const void* mask_aux, int ld_aux, const void* g_obj, const void* g_pred, const int* g_sidx,
                           const int* g_oidx, int g_din, int g_dp, int g_ldp, int g_nobj, const int* g_pidx, int g_npred,
                           int formats, void* workspace, size_t workspace_bytes, cudaStream_t stream, CsgReduceJob* job);
// csg_cast_bf16_multi with row pitches (elements; NULL / 0 = contiguous)
int csg_cast_bf16_multi_ld(int n, const void* const* src, void* const* dst, const int* rows, const int* cols,
                           const int* transpose, const int* ld_src, const int* ld_dst, int fp16, cudaStream_t stream);
// csg_colsum_bf16 without its final pass
int csg_colsum_bf16_deferred(const void* X, int M, int N, int ld, float* out, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream, CsgReduceJob* job);
// up to 4 column sums in one partial launch (job[i]: the final pass of matrix i)
int csg_colsum_bf16_multi_deferred(int n, const void* const* X, const int* M, const int* N, const int* ld, float* const* out,
                                   void* const* workspace, const size_t* workspace_bytes, cudaStream_t stream,
                                   CsgReduceJob* job);
// csg_triple_bwd_assemble_bf16 with the column sums of g (db2) and the per-predicate sums of the confidence gradient
// (d w_trans, graph.py:69-74) left as two jobs; dconf itself is not materialised.  workspace:
// csg_triple_bwd_assemble_bf16_deferred_workspace bytes.
size_t csg_triple_bwd_assemble_bf16_deferred_workspace(int NT, int H, int Dp, int P);
int csg_triple_bwd_assemble_bf16_deferred(const void* out, const float* dS, const void* d_newp, int ld_newp,
                                          const float* dcnt, const int* s_idx, const int* o_idx, const int* valid,
                                          const int* type32, const int* pred, const float* conf, const float* w_trans,
                                          int NT, int H, int Dp, int P, void* g, float* db2, float* dwt, int out_fp16,
                                          void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                          CsgReduceJob* job_db2, CsgReduceJob* job_dwt);
