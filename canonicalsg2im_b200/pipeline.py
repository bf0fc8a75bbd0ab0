"""One scene-graph -> layout training step on one GPU (the unit ``bench.py`` times):

    canonicalization (base_dataset.py:89-139)  ->  Sg2LayoutModel GCN stack + box_net (model.py:90-124)
    ->  boxes_to_layout canvas (layout.py:12-45, generator.py:81-96)  ->  backward  ->  [grad all-reduce]  ->  Adam

All heavy stages are csg2im kernels; torch provides memory, autograd bookkeeping, the tiny box loss and
the optimizer update.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import synth
from .canonicalize import canon_count_async, canon_emit, converse_tables
from .layout import layout_batched
from .model import Sg2LayoutModel, get_conv_converse, masked_box_loss
from .optim import FusedAdam
from .parallel import BucketedGradAllReduce


def build_opt(vocab: synth.Vocab, embedding_dim=128, gconv_dim=128, hidden_dim=512, num_layers=5):
    import argparse
    attrs = {"a%d" % i: {str(j): j for j in range(vocab.num_obj_classes if vocab.num_attributes == 1 else 8)}
             for i in range(vocab.num_attributes)}
    return argparse.Namespace(
        vocab={"attributes": attrs, "pred_idx_to_name": vocab.pred_names, "pred_name_to_idx": vocab.pred_ids},
        embedding_dim=embedding_dim, gconv_dim=gconv_dim, gconv_hidden_dim=hidden_dim, gconv_pooling="avg",
        gconv_num_layers=num_layers, mlp_normalization="none", mask_size=0, learned_init="uniform")


class HostBatch:
    """Flat host-side batch (pinned when CUDA is available): what a collate function would emit."""

    def __init__(self, graphs, seed=0, pin=True):
        self.B = len(graphs)
        self.tri_off = np.concatenate([[0], np.cumsum([len(g.triplets) for g in graphs])]).astype(np.int32)
        self.obj_off = np.concatenate([[0], np.cumsum([len(g.objs) for g in graphs])]).astype(np.int32)
        self.max_objs = max(len(g.objs) for g in graphs)
        arrs = dict(
            triplets=np.concatenate([g.triplets for g in graphs]).astype(np.int64),
            tri_off=self.tri_off, obj_off=self.obj_off,
            objs=np.concatenate([g.objs for g in graphs]).astype(np.int64),
            boxes=np.concatenate([g.boxes for g in graphs]).astype(np.float32),
            uniforms=synth.det_uniform(int(self.tri_off[-1]), seed * 7919 + 13),
        )
        self.t = {}
        for k, v in arrs.items():
            x = torch.from_numpy(np.ascontiguousarray(v))
            if pin and torch.cuda.is_available():
                x = x.pin_memory()
            self.t[k] = x
        self.nbytes = sum(x.numel() * x.element_size() for x in self.t.values())

    def to_device(self, device):
        d = {k: v.to(device, non_blocking=True) for k, v in self.t.items()}
        d["max_objs"] = self.max_objs
        d["B"] = self.B
        return d


class SgToLayoutStep:
    def __init__(self, vocab, device, precision="fp32", H=64, W=64, learned_converse=True,
                 learned_transitivity=True, lr=1e-4, seed=0, distributed=False):
        self.vocab, self.device, self.H, self.W = vocab, device, H, W
        self.flags = (learned_converse, learned_transitivity)
        self.model = Sg2LayoutModel(build_opt(vocab), precision=precision).to(device)
        st = {k: torch.from_numpy(v) for k, v in synth.make_state(vocab, seed=seed).items()}
        for i in range(len(self.model.gconvs)):
            st["gconvs.%d.predicates_transitive_weights" % i] = st["trans_candidates_weights"]
        self.model.load_state_dict(st, strict=True)
        self.refresh_tables()
        # gradient buckets in the order they complete: box_net, gconvs 4..0, then embeddings + shared weights
        layers = list(self.model.gconvs)
        buckets = [list(self.model.box_net.parameters())]
        for layer in reversed(layers):
            buckets.append([p for n, p in layer.named_parameters() if "predicates_transitive_weights" not in n])
        buckets.append(list(self.model.attribute_embedding.parameters()) + list(self.model.pred_embeddings.parameters())
                       + [self.model.trans_candidates_weights])
        self.reducer = BucketedGradAllReduce(buckets) if distributed else None
        params = [p for p in self.model.parameters() if p is not self.model.converse_candidates_weights]
        self.opt = FusedAdam(params, lr=lr)                      # torch.optim.Adam arithmetic, one multi-tensor launch

    def refresh_tables(self):
        """The reference pushes the symmetrised converse weights into the dataset every step
        (train.py:311-313); the CDF table is rebuilt from them here."""
        W = get_conv_converse(self.model).detach().double().cpu().numpy()
        cdf, vals = converse_tables(W, self.vocab.num_preds, self.vocab.meta_ids)
        self.tables = (torch.from_numpy(cdf).to(self.device), torch.from_numpy(vals).to(self.device))

    def prefetch(self, d):
        """Launch the counting pass of the canonicalization of batch ``d`` now (no host wait); ``step(d, ...)``
        later finds the output sizes on the host.  The reference canonicalizes in DataLoader workers ahead of the
        training step (base_dataset.py:89-139 inside ``__getitem__``); this is the same look-ahead on the device."""
        d["_canon_plan"] = canon_count_async(d["triplets"], d["tri_off"], d["obj_off"], self.vocab.num_preds,
                                             self.vocab.meta_ids, None, self.flags[0], self.flags[1], d["uniforms"],
                                             max_objs_per_graph=d["max_objs"], tables=self.tables)

    def canonicalize(self, d):
        if d.get("_canon_plan") is None:
            self.prefetch(d)
        return canon_emit(d.pop("_canon_plan"))

    def forward(self, d, res):
        obj_vecs, boxes_pred = self.model.forward_ragged(d["objs"], res.triplets, res.triplet_type, res.tri_off,
                                                         d["obj_off"])
        # canvas from GT boxes, as training does (train.py:358, generator.py:81-96); the __image__ dummy
        # has box -1 and contributes exact zeros, so no object filtering pass is needed
        canvas = layout_batched(obj_vecs, d["boxes"], d["obj_off"], self.H, self.W, max_objs_per_image=d["max_objs"])
        loss = masked_box_loss(boxes_pred, d["boxes"])                   # pix2pix_model.py:72-85
        return canvas, loss

    def step(self, d, canvas_grad, prefetch=None):
        """One training step on batch ``d``.  ``prefetch``: the batch of the NEXT step (may be ``d`` itself); its
        canonicalization counting pass is enqueued right behind this step's emit pass."""
        res = self.canonicalize(d)
        if prefetch is not None:
            self.prefetch(prefetch)
        canvas, loss = self.forward(d, res)
        torch.autograd.backward([canvas, loss], [canvas_grad, None])
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        if self.reducer is not None:
            self.reducer.zero()
        else:
            self.opt.zero_grad(set_to_none=True)
        return loss.detach(), int(res.triplets.shape[0])
