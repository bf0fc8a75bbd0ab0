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

from .graph import GraphTripleConv, TripleBatch, build_mlp, dense_mlp2, get_predicates_weights


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

    def forward(self, x):
        vecs = [self._modules["att_emb_%d" % k](x[..., k]) for k in range(x.size(-1))]
        v = torch.cat(vecs, dim=-1)
        if hasattr(self, "attribute_fc_gen"):
            v = self.attribute_fc_gen(v)
        return v


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
        obj_vecs = self.attribute_embedding(objs).reshape(B * O, -1)
        pred_vecs = self.pred_embeddings(triplets[:, :, 1]).reshape(B * T, -1)
        obj_vecs, boxes = self._run(batch, obj_vecs, pred_vecs)
        return obj_vecs.view(B, O, -1), boxes.view(B, O, 4), None

    def forward_ragged(self, objs, triplets, triplet_type, tri_off, obj_off):
        """Flat interface: objs [NO,A], triplets [NT,3] (graph-local ids), triplet_type [NT], offsets [B+1] i32
        -> (obj_vecs [NO,D], boxes_pred [NO,4])."""
        batch = TripleBatch.from_ragged(triplets, triplet_type, tri_off, obj_off, objs.size(0), self.padding_id)
        obj_vecs = self.attribute_embedding(objs)
        pred_vecs = self.pred_embeddings(triplets[:, 1])
        return self._run(batch, obj_vecs, pred_vecs)
