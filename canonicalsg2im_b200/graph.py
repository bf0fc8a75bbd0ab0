"""Triple graph convolution: drop-in for ``sg2im/graph.py`` of the reference.

``GraphTripleConv`` keeps the reference's constructor, ``forward`` signature,
parameter names and shapes (``sg2im/graph.py:22-24,44``; state-dict keys
``net1.0.weight`` ... ``predicates_transitive_weights``), so checkpoints load
unchanged.  Internally the padded ``[B, O, .]`` / ``[B, T, .]`` batch is run as
a flat batch (padded rows are ordinary rows with ``valid = 0``), the gather of
subject/object rows is fused into the first net1 GEMM, pooling is a
deterministic CSR segmented reduction, and fwd + bwd are hand-written kernels
behind one ``torch.autograd.Function`` per layer.

``GraphTripleConvNet`` is the 5-layer stack of ``sg2im/model.py:111-112`` on
flat (ragged) tensors: no padded rows are computed at all.

``precision``: ``"fp32"`` = fp32 FMA GEMMs (1e-5 parity with the reference);
``"bf16"`` = tcgen05 tensor-core GEMMs with bf16 operands / fp32 accumulation
(1e-2 tolerance, the throughput path); ``"fp16"`` = the same engine with fp16
forward tensors (8x finer rounding at the same tensor-pipe rate), inference only.
"""
import torch
import torch.nn as nn

from . import _lib
from . import ops
from .ops import lib, ptr, need_cuda, f32c, workspace, _stream, Gather
from .ops import A_ROW, A_COL, A_GATHER, B_NK, B_KN, B_GATHER

ORIGINAL_EDGE, TRANSITIVE_EDGE, SYMMETRIC_EDGE, ANTI_SYMMETRIC_EDGE = 0, 1, 2, 3   # base_dataset.py:7-10


# --------------------------------------------------------------------------------------------
# per-batch index structures (shared by all layers, forward and backward)
# --------------------------------------------------------------------------------------------
class TripleBatch:
    """Flat triple batch: int32 global subject/object ids, predicate ids, edge types, valid flags
    and the two stable CSR orderings (by subject / by object)."""

    def __init__(self, s_idx, o_idx, pred, type32, valid, tri_off, obj_off, NO):
        self.s_idx, self.o_idx, self.pred, self.type32, self.valid = s_idx, o_idx, pred, type32, valid
        self.tri_off, self.obj_off = tri_off, obj_off
        self.NO, self.NT, self.B = int(NO), int(s_idx.numel()), int(tri_off.numel() - 1)
        dev = s_idx.device
        self.rowptr_s = torch.empty(self.NO + 1, dtype=torch.int32, device=dev)
        self.rowptr_o = torch.empty(self.NO + 1, dtype=torch.int32, device=dev)
        self.perm_s = torch.empty(max(self.NT, 1), dtype=torch.int32, device=dev)
        self.perm_o = torch.empty(max(self.NT, 1), dtype=torch.int32, device=dev)
        L = lib()
        ws = workspace(L.csg_csr_workspace(self.NO), dev)
        rc = L.csg_csr_build(ptr(s_idx), ptr(o_idx), ptr(tri_off), ptr(obj_off), self.B, self.NT, self.NO,
                             ptr(self.rowptr_s), ptr(self.perm_s), ptr(self.rowptr_o), ptr(self.perm_o),
                             ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "csg_csr_build")

    def index_array(self):
        """HOST array of the 9 device pointers csg_gconv_bf16_fwd/bwd take (include/csg2im.h)."""
        arr = getattr(self, "_index_array", None)
        if arr is None:
            import ctypes
            ts = (self.s_idx, self.o_idx, self.pred, self.type32, self.valid, self.rowptr_s, self.perm_s,
                  self.rowptr_o, self.perm_o)
            arr = self._index_array = (ctypes.c_void_p * 9)(*[t.data_ptr() for t in ts])
        return arr

    # ---- constructors
    @staticmethod
    def _alloc(NT, dev):
        return [torch.empty(max(NT, 1), dtype=torch.int32, device=dev) for _ in range(5)]

    @staticmethod
    def _uniform_off(B, stride, dev):
        off = torch.empty(B + 1, dtype=torch.int32, device=dev)
        _lib.check(lib().csg_offsets_uniform(ptr(off), B, stride, _stream()), "csg_offsets_uniform")
        return off

    @classmethod
    def from_padded_edges(cls, edges, pred_indicators, triplet_type, predicate_ids, O, num_preds=0):
        """The argument layout of ``GraphTripleConv.forward`` (graph.py:44): edges [B,T,2] i64,
        pred_indicators [B,T] bool, triplet_type [B,T] i64, predicate_ids [B,T] i64.  ``num_preds`` > 0 also
        range-checks the predicate ids (they index the transitive weights, graph.py:73)."""
        need_cuda(edges, pred_indicators, triplet_type, predicate_ids)
        B, T = edges.shape[0], edges.shape[1]
        dev = edges.device
        NT = B * T
        s_idx, o_idx, pred, type32, valid = cls._alloc(NT, dev)
        e = edges.contiguous().to(torch.int64)
        pid = predicate_ids.contiguous().to(torch.int64)
        ind = pred_indicators.contiguous().to(torch.uint8) if pred_indicators is not None else None
        tt = triplet_type.contiguous().to(torch.int64) if triplet_type is not None else None
        rc = lib().csg_triple_prep_edges(ptr(e), ptr(pid), ptr(ind), ptr(tt), 0, 0, B, NT, T, O, int(num_preds),
                                         ptr(s_idx), ptr(o_idx), ptr(pred), ptr(type32), ptr(valid), _stream())
        _lib.check(rc, "csg_triple_prep_edges")
        return cls(s_idx[:NT], o_idx[:NT], pred[:NT], type32[:NT], valid[:NT],
                   cls._uniform_off(B, T, dev), cls._uniform_off(B, O, dev), B * O)

    @classmethod
    def from_padded_triplets(cls, triplets, triplet_type, padding_id, O, num_preds=0):
        """``Sg2LayoutModel.forward``'s layout (model.py:104-107): triplets [B,T,3] i64."""
        need_cuda(triplets, triplet_type)
        B, T = triplets.shape[0], triplets.shape[1]
        dev = triplets.device
        NT = B * T
        s_idx, o_idx, pred, type32, valid = cls._alloc(NT, dev)
        tr = triplets.contiguous().to(torch.int64)
        tt = triplet_type.contiguous().to(torch.int64) if triplet_type is not None else None
        rc = lib().csg_triple_prep(ptr(tr), ptr(tt), 0, 0, B, NT, T, O, int(padding_id), int(num_preds),
                                   ptr(s_idx), ptr(o_idx), ptr(pred), ptr(type32), ptr(valid), _stream())
        _lib.check(rc, "csg_triple_prep")
        return cls(s_idx[:NT], o_idx[:NT], pred[:NT], type32[:NT], valid[:NT],
                   cls._uniform_off(B, T, dev), cls._uniform_off(B, O, dev), B * O)

    @classmethod
    def from_ragged(cls, triplets, triplet_type, tri_off, obj_off, NO, padding_id=-1, num_preds=0):
        """Flat batch: triplets [NT,3] i64 with per-graph LOCAL object ids, tri_off/obj_off [B+1] i32."""
        need_cuda(triplets, triplet_type, tri_off, obj_off)
        NT = triplets.shape[0]
        dev = triplets.device
        B = tri_off.numel() - 1
        s_idx, o_idx, pred, type32, valid = cls._alloc(NT, dev)
        tr = triplets.contiguous().to(torch.int64)
        tt = triplet_type.contiguous().to(torch.int64) if triplet_type is not None else None
        tri_off = tri_off.to(torch.int32).contiguous()
        obj_off = obj_off.to(torch.int32).contiguous()
        rc = lib().csg_triple_prep(ptr(tr), ptr(tt), ptr(tri_off), ptr(obj_off), B, NT, 0, 0, int(padding_id), int(num_preds),
                                   ptr(s_idx), ptr(o_idx), ptr(pred), ptr(type32), ptr(valid), _stream())
        _lib.check(rc, "csg_triple_prep")
        return cls(s_idx[:NT], o_idx[:NT], pred[:NT], type32[:NT], valid[:NT], tri_off, obj_off, NO)


def triple_confidence(batch, w_trans):
    """graph.py:69-74 -> conf [NT] fp32."""
    conf = torch.empty(max(batch.NT, 1), dtype=torch.float32, device=w_trans.device)
    w = f32c(w_trans)
    _lib.check(lib().csg_triple_conf(ptr(batch.type32), ptr(batch.pred), ptr(w), batch.NT, ptr(conf), _stream()),
               "csg_triple_conf")
    return conf[:batch.NT]


def segpool(X, col_s, col_o, W, batch, conf=None, avg=True):
    """Segmented reduction of triple rows onto objects (graph.py:85-107 when ``avg``)."""
    out = torch.empty((batch.NO, W), dtype=torch.float32, device=X.device)
    cnt = torch.empty(batch.NO, dtype=torch.float32, device=X.device) if avg else None
    rc = lib().csg_segpool_f32(ptr(X), X.stride(0), col_s, col_o, W, ptr(batch.rowptr_s), ptr(batch.perm_s),
                               ptr(batch.rowptr_o), ptr(batch.perm_o), ptr(batch.valid) if avg else 0,
                               ptr(conf) if avg else 0, batch.NO, ptr(out), out.stride(0), ptr(cnt), int(avg),
                               _stream())
    _lib.check(rc, "csg_segpool_f32")
    return out, cnt


# --------------------------------------------------------------------------------------------
# one layer, fp32 engine
# --------------------------------------------------------------------------------------------
class _TripleConvF32(torch.autograd.Function):
    """forward(obj [NO,Din], pred [NT,Dp], w1,b1,w2,b2,w3,b3,w4,b4, w_trans) -> (new_obj [NO,Dout], new_p [NT,Dpo])"""

    @staticmethod
    def forward(ctx, batch, H, Dpo, obj, pred, w1, b1, w2, b2, w3, b3, w4, b4, w_trans):
        NT, NO = batch.NT, batch.NO
        obj = f32c(obj)
        pred = pred.float() if pred.dtype != torch.float32 else pred
        if pred.stride(-1) != 1 or pred.stride(0) % 4 != 0:
            pred = pred.contiguous()
        w1, w2, w3, w4 = f32c(w1), f32c(w2), f32c(w3), f32c(w4)
        g = Gather(obj, pred, batch.s_idx, batch.o_idx)
        conf = triple_confidence(batch, w_trans)
        hidden = ops.gemm_f32(A_GATHER, B_NK, NT, H, g.width, None, w1, bias=f32c(b1), relu=True, gather=g, lda=0)
        out = ops.gemm_f32(A_ROW, B_NK, NT, 2 * H + Dpo, H, hidden, w2, bias=f32c(b2), relu=True, rowscale=conf)
        pooled, cnt = segpool(out, 0, H + Dpo, H, batch, conf, avg=True)
        h2 = ops.gemm_f32(A_ROW, B_NK, NO, H, H, pooled, w3, bias=f32c(b3), relu=True)
        new_obj = ops.gemm_f32(A_ROW, B_NK, NO, w4.shape[0], H, h2, w4, bias=f32c(b4), relu=True)
        new_p = out[:, H:H + Dpo]
        ctx.batch, ctx.H, ctx.Dpo = batch, H, Dpo
        ctx.save_for_backward(obj, pred, w1, w2, w3, w4, f32c(w_trans), conf, hidden, out, pooled, cnt, h2, new_obj)
        ctx.set_materialize_grads(False)
        return new_obj, new_p

    @staticmethod
    def backward(ctx, d_obj_out, d_newp):
        batch, H, Dpo = ctx.batch, ctx.H, ctx.Dpo
        obj, pred, w1, w2, w3, w4, w_trans, conf, hidden, out, pooled, cnt, h2, new_obj = ctx.saved_tensors
        NT, NO = batch.NT, batch.NO
        dev = obj.device
        L = lib()
        Dout = w4.shape[0]
        Din, Dp = obj.shape[1], pred.shape[1]
        if d_obj_out is None:
            d_obj_out = torch.zeros((NO, Dout), dtype=torch.float32, device=dev)
        # ---- net2 backward (graph.py:109-110)
        g4 = ops.relu_mask_f32(d_obj_out, new_obj)
        dw4 = ops.gemm_f32(A_COL, B_KN, Dout, H, NO, g4, h2)
        db4 = ops.colsum_f32(g4)
        dh2 = ops.gemm_f32(A_ROW, B_KN, NO, H, Dout, g4, w4, mask_aux=h2)
        dw3 = ops.gemm_f32(A_COL, B_KN, H, H, NO, dh2, pooled)
        db3 = ops.colsum_f32(dh2)
        dpooled = ops.gemm_f32(A_ROW, B_KN, NO, H, H, dh2, w3)
        # ---- pooling backward (graph.py:85-107)
        dS = torch.empty_like(dpooled)
        dcnt = torch.empty(NO, dtype=torch.float32, device=dev)
        _lib.check(L.csg_pool_bwd_obj(ptr(dpooled), ptr(pooled), ptr(cnt), NO, H, ptr(dS), ptr(dcnt), _stream()),
                   "csg_pool_bwd_obj")
        dnp = f32c(d_newp) if d_newp is not None else None
        Wd = 2 * H + Dpo
        g = torch.empty((NT, Wd), dtype=torch.float32, device=dev)
        dconf = torch.empty(max(NT, 1), dtype=torch.float32, device=dev)
        rc = L.csg_triple_bwd_assemble(ptr(out), ptr(dS), ptr(dnp), ptr(dcnt), ptr(batch.s_idx), ptr(batch.o_idx),
                                       ptr(batch.valid), ptr(batch.type32), ptr(conf), NT, H, Dpo, ptr(g), ptr(dconf),
                                       _stream())
        _lib.check(rc, "csg_triple_bwd_assemble")
        # ---- net1 backward (graph.py:66-67)
        dw2 = ops.gemm_f32(A_COL, B_KN, Wd, H, NT, g, hidden)
        db2 = ops.colsum_f32(g)
        dhid = ops.gemm_f32(A_ROW, B_KN, NT, H, Wd, g, w2, mask_aux=hidden)
        gat = Gather(obj, pred, batch.s_idx, batch.o_idx)
        dw1 = ops.gemm_f32(A_COL, B_GATHER, H, gat.width, NT, dhid, None, gather=gat, ldb=0)
        db1 = ops.colsum_f32(dhid)
        dX = ops.gemm_f32(A_ROW, B_KN, NT, gat.width, H, dhid, w1)
        # ---- gather backward (graph.py:63-64): segmented sums over ALL triples
        dobj, _ = segpool(dX, 0, Din + Dp, Din, batch, avg=False)
        dpred = dX[:, Din:Din + Dp]
        # ---- confidence backward (graph.py:69-74)
        P = w_trans.numel()
        dwt = torch.empty(P, dtype=torch.float32, device=dev)
        ws = workspace(L.csg_conf_bwd_workspace(P), dev)
        rc = L.csg_conf_bwd(ptr(dconf), ptr(batch.type32), ptr(batch.pred), ptr(w_trans), NT, P, ptr(dwt),
                            ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "csg_conf_bwd")
        return None, None, None, dobj, dpred, dw1, db1, dw2, db2, dw3, db3, dw4, db4, dwt


class _DenseMLP2F32(torch.autograd.Function):
    """y = [relu](relu(x W0^T + b0) W1^T + b1)  (layers.py:6-25 with two Linear layers; box_net, model.py:58-60)."""

    @staticmethod
    def forward(ctx, x, w0, b0, w1, b1, final_relu):
        x, w0, w1 = f32c(x), f32c(w0), f32c(w1)
        M = x.shape[0]
        h = ops.gemm_f32(A_ROW, B_NK, M, w0.shape[0], w0.shape[1], x, w0, bias=f32c(b0), relu=True)
        y = ops.gemm_f32(A_ROW, B_NK, M, w1.shape[0], w1.shape[1], h, w1, bias=f32c(b1), relu=bool(final_relu))
        ctx.save_for_backward(x, w0, w1, h, y)
        ctx.final_relu = bool(final_relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w0, w1, h, y = ctx.saved_tensors
        M = x.shape[0]
        dy = ops.relu_mask_f32(dy, y) if ctx.final_relu else f32c(dy)
        N1, K1 = w1.shape
        N0, K0 = w0.shape
        # box_net's output width (4) is below the 4-float alignment of the transposed loads: pad the lda
        if N1 % 4 != 0:
            pad = torch.zeros((M, (N1 + 3) // 4 * 4), dtype=torch.float32, device=dy.device)
            pad[:, :N1] = dy
            dy_k = pad
            w1_k = torch.zeros((dy_k.shape[1], K1), dtype=torch.float32, device=dy.device)
            w1_k[:N1] = w1
        else:
            dy_k, w1_k = dy, w1
        dw1 = ops.gemm_f32(A_COL, B_KN, N1, K1, M, dy_k, h)
        db1 = ops.colsum_f32(dy)
        dh = ops.gemm_f32(A_ROW, B_KN, M, K1, dy_k.shape[1], dy_k, w1_k, mask_aux=h)
        dw0 = ops.gemm_f32(A_COL, B_KN, N0, K0, M, dh, x)
        db0 = ops.colsum_f32(dh)
        dx = ops.gemm_f32(A_ROW, B_KN, M, K0, N0, dh, w0)
        return dx, dw0, db0, dw1, db1, None


class _LinearF32(torch.autograd.Function):
    """y = x W^T + b on the fp32 engine (``attribute_fc_gen``, attribute_embed.py:24-25,46-47)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = f32c(x), f32c(w)
        y = ops.gemm_f32(A_ROW, B_NK, x.shape[0], w.shape[0], w.shape[1], x, w, bias=f32c(b))
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = f32c(dy)
        M, (N, K) = x.shape[0], w.shape
        dw = ops.gemm_f32(A_COL, B_KN, N, K, M, dy, x)
        db = ops.colsum_f32(dy)
        dx = ops.gemm_f32(A_ROW, B_KN, M, K, N, dy, w) if ctx.needs_input_grad[0] else None
        return dx, dw, db


def linear(x, w, b, precision="fp32"):
    """``F.linear`` on the csg2im GEMMs (2-D ``x``).  fp32: feature widths must be multiples of 4;
    bf16 (tcgen05, bf16 output) needs multiples of 64 and falls back to the fp32 engine otherwise."""
    need_cuda(x, w, b)
    if x.dim() != 2:
        raise ValueError("linear expects a 2-D input (got %s)" % (tuple(x.shape),))
    if precision in ("bf16", "fp16") and w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0:
        from . import graph_tc
        return graph_tc._LinearBF16.apply(torch.is_grad_enabled(), x, w, b)
    if w.shape[0] % 4 or w.shape[1] % 4:
        raise _lib.CsgError("linear: feature widths must be multiples of 4 (got %s)" % (tuple(w.shape),))
    return _LinearF32.apply(x, w, b)


def dense_mlp2(x, w0, b0, w1, b1, final_relu, precision="fp32"):
    need_cuda(x, w0, w1)
    if precision in ("bf16", "fp16"):
        from . import graph_tc
        return graph_tc.dense_mlp2(x, w0, b0, w1, b1, final_relu)
    return _DenseMLP2F32.apply(x, w0, b0, w1, b1, final_relu)


def triple_conv(batch, obj, pred, params, w_trans, hidden_dim, pred_out_dim, precision="fp32"):
    """One layer on flat tensors.  params = (w1, b1, w2, b2, w3, b3, w4, b4)."""
    need_cuda(obj, pred, w_trans)
    if precision in ("bf16", "fp16"):
        if precision == "fp16" and obj.dtype != torch.float16:      # the layer called on its own (GraphTripleConv.forward)
            obj, pred = obj.to(torch.float16), pred.to(torch.float16)
        from . import graph_tc
        return graph_tc.triple_conv(batch, obj, pred, params, w_trans, hidden_dim, pred_out_dim)
    return _TripleConvF32.apply(batch, hidden_dim, pred_out_dim, obj, pred, *params, w_trans)


# --------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------
def _init_weights(module):                       # graph.py:11-14
    if hasattr(module, "weight") and isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)


def build_mlp(dim_list, final_nonlinearity="relu"):
    """Parameter container with the reference's nn.Sequential indices (layers.py:6-25, batch_norm='none',
    dropout=0): Linear at 0, ReLU at 1, Linear at 2, [ReLU at 3].  Never called as a module."""
    layers = []
    for i in range(len(dim_list) - 1):
        layers.append(nn.Linear(dim_list[i], dim_list[i + 1]))
        if i != len(dim_list) - 2:
            layers.append(nn.ReLU())
    if final_nonlinearity is not None:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


def _tensor_key(x):
    return None if x is None else (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), x.dtype)


class _BatchCache:
    """All layers of a forward pass receive the same index tensors (model.py:111-112): the CSR is built once.
    One entry per (device, stream), keyed on storage + version so in-place edits invalidate it; the entry keeps
    the keyed tensors alive, so a data_ptr cannot be recycled while it exists, and is dropped when a different
    batch arrives on that stream (``clear()`` releases everything)."""

    def __init__(self):
        self.entries = {}

    def get(self, edges, pred_indicators, triplet_type, predicate_ids, O, num_preds):
        need_cuda(edges, pred_indicators, triplet_type, predicate_ids)
        slot = (edges.device, torch.cuda.current_stream(edges.device).cuda_stream)
        key = (_tensor_key(edges), _tensor_key(pred_indicators), _tensor_key(triplet_type), _tensor_key(predicate_ids),
               O, num_preds)
        hit = self.entries.get(slot)
        if hit is not None and hit[0] == key:
            return hit[1]
        b = TripleBatch.from_padded_edges(edges, pred_indicators, triplet_type, predicate_ids, O, num_preds)
        self.entries[slot] = (key, b, (edges, pred_indicators, triplet_type, predicate_ids))
        return b

    def clear(self):
        self.entries.clear()


_BATCH_CACHE = _BatchCache()


def _cached_batch(edges, pred_indicators, triplet_type, predicate_ids, O, num_preds=0):
    return _BATCH_CACHE.get(edges, pred_indicators, triplet_type, predicate_ids, O, num_preds)


class GraphTripleConv(nn.Module):
    """A single layer of scene graph convolution (sg2im/graph.py:17-113)."""

    def __init__(self, obj_input_dim, object_output_dim, predicate_input_dim, predicate_output_dim, hidden_dim,
                 num_attributes, pooling="avg", mlp_normalization="none", predicates_transitive_weights=None,
                 return_new_p_vecs=True, precision="fp32"):
        super().__init__()
        assert pooling in ["sum", "avg"], 'Invalid pooling "%s"' % pooling          # graph.py:31
        if mlp_normalization != "none":
            raise ValueError("only mlp_normalization='none' (the reference default, args.py:53) is supported")
        self.return_new_p_vecs = return_new_p_vecs
        self.hidden_dim = hidden_dim
        self.obj_input_dim, self.predicate_input_dim = obj_input_dim, predicate_input_dim
        self.num_attributes = num_attributes
        self.predicate_output_dim = predicate_output_dim
        self.pooling = pooling            # accepted and ignored, as in the reference
        self.precision = precision
        self.net1 = build_mlp([2 * obj_input_dim + predicate_input_dim, hidden_dim,
                               2 * hidden_dim + predicate_output_dim])
        self.net1.apply(_init_weights)
        self.net2 = build_mlp([hidden_dim, hidden_dim, object_output_dim])
        self.net2.apply(_init_weights)
        self.predicates_transitive_weights = predicates_transitive_weights           # graph.py:42

    def layer_params(self):
        return (self.net1[0].weight, self.net1[0].bias, self.net1[2].weight, self.net1[2].bias,
                self.net2[0].weight, self.net2[0].bias, self.net2[2].weight, self.net2[2].bias)

    def forward_flat(self, batch, obj_vecs, pred_vecs):
        new_obj, new_p = triple_conv(batch, obj_vecs, pred_vecs, self.layer_params(),
                                     self.predicates_transitive_weights, self.hidden_dim,
                                     self.predicate_output_dim, self.precision)
        return new_obj, (new_p if self.return_new_p_vecs else pred_vecs)

    def forward(self, obj_vecs, pred_vecs, edges, pred_indicators, triplet_type, predicate_ids):
        """Padded interface of the reference (graph.py:44-113):
        obj_vecs [B,O,Din], pred_vecs [B,T,Dp], edges [B,T,2] i64, pred_indicators [B,T] bool,
        triplet_type [B,T] i64, predicate_ids [B,T] i64 -> (new_obj [B,O,Dout], new_p [B,T,Dp_out])."""
        B, O, T = obj_vecs.size(0), obj_vecs.size(1), pred_vecs.size(1)
        w = self.predicates_transitive_weights
        batch = _cached_batch(edges, pred_indicators, triplet_type, predicate_ids, O, w.numel() if w is not None else 0)
        new_obj, new_p = self.forward_flat(batch, obj_vecs.reshape(B * O, -1), pred_vecs.reshape(B * T, -1))
        return new_obj.view(B, O, -1), new_p.reshape(B, T, -1)


class GraphTripleConvNet(nn.Module):
    """The GraphTripleConv stack of ``sg2im/model.py:35-55,111-112`` on flat tensors.

    forward(obj_vecs [NO, D0], pred_vecs [NT, Dp0], batch: TripleBatch) -> (obj_vecs [NO, D], pred_vecs [NT, D])"""

    def __init__(self, obj_input_dim, pred_input_dim, gconv_dim=128, hidden_dim=512, num_layers=5,
                 num_attributes=1, predicates_transitive_weights=None, precision="fp32"):
        super().__init__()
        self.gconvs = nn.ModuleList()
        d_obj, d_pred = obj_input_dim, pred_input_dim
        for _ in range(num_layers):
            self.gconvs.append(GraphTripleConv(d_obj, gconv_dim, d_pred, gconv_dim, hidden_dim, num_attributes,
                                               predicates_transitive_weights=predicates_transitive_weights,
                                               precision=precision))
            d_obj, d_pred = gconv_dim, gconv_dim

    def forward(self, obj_vecs, pred_vecs, batch):
        for layer in self.gconvs:
            obj_vecs, pred_vecs = layer.forward_flat(batch, obj_vecs, pred_vecs)
        return obj_vecs, pred_vecs


def get_predicates_weights(num_preds, learned_init):
    """sg2im/graph.py:115-127."""
    if learned_init == "uniform":
        w = torch.nn.Parameter(torch.zeros(num_preds), requires_grad=True)
        w.data.uniform_(-1, 1)
    elif learned_init == "-4":
        w = torch.nn.Parameter(-4 * torch.ones(num_preds), requires_grad=True)
    elif learned_init == "0":
        w = torch.nn.Parameter(torch.zeros(num_preds), requires_grad=True)
    elif learned_init == "4":
        w = torch.nn.Parameter(4 * torch.ones(num_preds), requires_grad=True)
    else:
        raise ValueError()
    return w
