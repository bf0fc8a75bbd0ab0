"""Scene-graph -> layout model: drop-in for ``sg2im/model.py::Sg2LayoutModel`` (reference
``sg2im/model.py:18-124``) built on the csg2im kernels.

State-dict keys equal the reference's (``attribute_embedding.att_emb_{k}.weight``,
``pred_embeddings.weight``, ``trans_candidates_weights``, ``converse_candidates_weights``,
``gconvs.{i}.net1.{0,2}.*``, ``gconvs.{i}.net2.{0,2}.*``, ``gconvs.{i}.predicates_transitive_weights``,
``box_net.{0,2}.*``), so reference checkpoints load with ``strict=True`` when ``mask_size == 0``.
The optional cuDNN ``mask_net`` conv stack (``model.py:62-75``) is outside the hot path and not built.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .graph import GraphTripleConv, TripleBatch, build_mlp, dense_mlp2, get_predicates_weights
from .ops import lib, ptr, need_cuda, workspace, _stream


class _EmbedFn(torch.autograd.Function):
    """``nn.Embedding`` lookup (attribute_embed.py:38-48, model.py:109) on csg_embed_fwd / csg_embed_bwd:
    rows are written straight in the GCN's operand type and the table gradient is a deterministic segmented sum."""

    @staticmethod
    def forward(ctx, table, idx, out_dtype):
        need_cuda(table, idx)
        if idx.dtype != torch.int64:
            idx = idx.to(torch.int64)
        if idx.dim() != 1:
            raise ValueError("embedding index must be 1-D (got %s)" % (tuple(idx.shape),))
        if idx.numel() and idx.stride(0) < 1:
            idx = idx.contiguous()
        tab = table.detach()
        if tab.dtype != torch.float32 or not tab.is_contiguous():
            tab = tab.float().contiguous()
        n, (V, E) = idx.numel(), tab.shape
        out = torch.empty((n, E), dtype=out_dtype, device=tab.device)
        rc = lib().csg_embed_fwd(ptr(tab), ptr(idx), idx.stride(0) if n else 1, n, V, E, ptr(out), E,
                                 int(out_dtype == torch.bfloat16), _stream())
        _lib.check(rc, "csg_embed_fwd")
        ctx.save_for_backward(idx)
        ctx.dims = (n, V, E)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        n, V, E = ctx.dims
        if dout.dtype not in (torch.float32, torch.bfloat16):
            dout = dout.float()
        if dout.stride(-1) != 1 or dout.stride(0) % 4 != 0 or dout.data_ptr() % 16 != 0:
            dout = dout.contiguous()
        L = lib()
        if (dout.dtype == torch.bfloat16 and n >= 1024 and E % 32 == 0 and dout.stride(0) % 8 == 0
                and dout.data_ptr() % 16 == 0):
            # large bf16 lookups (the predicate rows of the tensor-core engine): dtable = onehot^T dout on tcgen05
            from .ops import gemm_bf16
            ld = (V + 63) // 64 * 64
            onehot = torch.empty((n, ld), dtype=torch.bfloat16, device=dout.device)
            _lib.check(L.csg_onehot_bf16(ptr(idx), idx.stride(0), n, V, ptr(onehot), ld, _stream()), "csg_onehot_bf16")
            return gemm_bf16(ld, E, n, onehot, dout, mn_major=True)[:V], None, None
        dtable = torch.empty((V, E), dtype=torch.float32, device=dout.device)
        ws = workspace(L.csg_embed_bwd_workspace(n, V, E), dout.device)
        rc = L.csg_embed_bwd(ptr(dout), dout.stride(0) if n else E, int(dout.dtype == torch.bfloat16), ptr(idx),
                             idx.stride(0) if n else 1, n, V, E, ptr(dtable), ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "csg_embed_bwd")
        return dtable, None, None


def embedding_lookup(weight, idx, out_dtype=torch.float32):
    """``weight[idx]`` for a 1-D int64 index (any stride); ids must lie in [0, weight.shape[0])."""
    return _EmbedFn.apply(weight, idx, out_dtype)


class _BoxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        need_cuda(pred, gt)
        p, g = pred.detach().float().contiguous(), gt.float().contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        dpred = torch.empty_like(p)
        _lib.check(lib().csg_box_loss(ptr(p), ptr(g), p.shape[0], ptr(loss), ptr(dpred), _stream()), "csg_box_loss")
        ctx.save_for_backward(dpred)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        (dpred,) = ctx.saved_tensors
        return dpred * dloss, None


def masked_box_loss(boxes_pred, boxes_gt):
    """Mean smooth-L1 between predicted and ground-truth boxes over the REAL objects (gt >= 0; the ``__image__``
    dummy has box -1), the regression term of pix2pix_model.py:72-85 on a flat batch.  One launch, no host sync."""
    return _BoxLossFn.apply(boxes_pred.reshape(-1, 4), boxes_gt.reshape(-1, 4))


def get_conv_converse(model):
    """sg2im/model.py:8-15 — symmetrised converse weights (triu + triu^T)."""
    if isinstance(model, dict):
        base = model["sg_to_layout.module.converse_candidates_weights"]
    elif hasattr(model, "sg_to_layout"):
        base = model.sg_to_layout.module.converse_candidates_weights
    else:
        base = model.converse_candidates_weights
    triu = torch.triu(base, diagonal=0)
    return triu + triu.t()


class AttributeEmbeddings(nn.Module):
    """sg2im/attribute_embed.py:18-48 (lookup tables feeding the GCN; plain torch gathers)."""

    def __init__(self, attributes, embedding_dim, use_attr_fc_gen=False):
        super().__init__()
        num_attr = len(attributes)
        if num_attr > 1 or use_attr_fc_gen:
            self.attribute_fc_gen = nn.Linear(num_attr * embedding_dim, num_attr * embedding_dim)
        for i, name in enumerate(list(attributes)):
            self.add_module("att_emb_%d" % i, nn.Embedding(max(attributes[name].values()) + 1, embedding_dim))
        self.num_attr = num_attr

    def forward(self, x, out_dtype=torch.float32):
        lead = x.shape[:-1]
        flat = x.reshape(-1, x.size(-1))
        if not hasattr(self, "attribute_fc_gen") and flat.size(-1) == 1:
            # single attribute (COCO / VG): the rows go straight out in the GCN's operand type
            v = embedding_lookup(self._modules["att_emb_0"].weight, flat[:, 0], out_dtype)
            return v.view(*lead, -1)
        vecs = [embedding_lookup(self._modules["att_emb_%d" % k].weight, flat[:, k]) for k in range(flat.size(-1))]
        v = torch.cat(vecs, dim=-1)
        if hasattr(self, "attribute_fc_gen"):
            v = self.attribute_fc_gen(v)
        return v.view(*lead, -1)


class Sg2LayoutModel(nn.Module):
    def __init__(self, opt, precision="fp32"):
        super().__init__()
        args = vars(opt) if not isinstance(opt, dict) else dict(opt)
        self.args = args
        self.vocab = args["vocab"]
        self.precision = precision
        emb = args["embedding_dim"]
        self.attribute_embedding = AttributeEmbeddings(self.vocab["attributes"], emb)
        num_preds = len(self.vocab["pred_idx_to_name"])
        self.pred_embeddings = nn.Embedding(num_preds, emb)
        num_attributes = len(self.vocab["attributes"].keys())
        init = args.get("learned_init", "uniform")
        self.trans_candidates_weights = get_predicates_weights(num_preds, init)
        self.converse_candidates_weights = get_predicates_weights((num_preds, num_preds), init)
        if (args.get("mask_size") or 0) > 0:
            raise NotImplementedError("mask_net (cuDNN conv stack, model.py:62-75) is outside the csg2im hot path")
        gdim, hdim = args["gconv_dim"], args["gconv_hidden_dim"]
        self.gconvs = nn.ModuleList()
        d_obj, d_pred = num_attributes * emb, emb
        for _ in range(args["gconv_num_layers"]):
            self.gconvs.append(GraphTripleConv(d_obj, gdim, d_pred, gdim, hdim, num_attributes,
                                               pooling=args.get("gconv_pooling", "avg"),
                                               mlp_normalization=args.get("mlp_normalization", "none"),
                                               predicates_transitive_weights=self.trans_candidates_weights,
                                               precision=precision))
            d_obj, d_pred = gdim, gdim
        self.box_net = build_mlp([gdim, hdim, 4], final_nonlinearity=None)      # model.py:58-60
        self.mask_net = None
        self.padding_id = self.vocab["pred_name_to_idx"]["__padding__"]

    def _act_dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    def _run(self, batch, obj_vecs, pred_vecs):
        for layer in self.gconvs:
            obj_vecs, pred_vecs = layer.forward_flat(batch, obj_vecs, pred_vecs)
        boxes = dense_mlp2(obj_vecs, self.box_net[0].weight, self.box_net[0].bias,
                           self.box_net[2].weight, self.box_net[2].bias, False, self.precision)
        return obj_vecs, boxes

    def forward(self, objs, triplets, triplet_type, boxes_gt=None, masks_gt=None):
        """Padded interface (model.py:90-124): objs [B,O,A] i64, triplets [B,T,3] i64, triplet_type [B,T] i64
        -> (obj_vecs [B,O,D], boxes_pred [B,O,4], None)."""
        B, O, T = objs.size(0), objs.size(1), triplets.size(1)
        batch = TripleBatch.from_padded_triplets(triplets, triplet_type, self.padding_id, O)
        obj_vecs = self.attribute_embedding(objs, self._act_dtype()).reshape(B * O, -1)
        pred_vecs = embedding_lookup(self.pred_embeddings.weight, triplets.reshape(B * T, 3)[:, 1], self._act_dtype())
        obj_vecs, boxes = self._run(batch, obj_vecs, pred_vecs)
        return obj_vecs.view(B, O, -1), boxes.view(B, O, 4), None

    def forward_ragged(self, objs, triplets, triplet_type, tri_off, obj_off):
        """Flat interface: objs [NO,A], triplets [NT,3] (graph-local ids), triplet_type [NT], offsets [B+1] i32
        -> (obj_vecs [NO,D], boxes_pred [NO,4])."""
        batch = TripleBatch.from_ragged(triplets, triplet_type, tri_off, obj_off, objs.size(0), self.padding_id)
        obj_vecs = self.attribute_embedding(objs, self._act_dtype())
        pred_vecs = embedding_lookup(self.pred_embeddings.weight, triplets[:, 1], self._act_dtype())
        return self._run(batch, obj_vecs, pred_vecs)
