"""CPU restatement of one scene-graph -> layout training step, assembled from the oracle pieces the way
the reference's training loop runs them (TEST INFRASTRUCTURE / CPU baseline only):

    Dataset.__getitem__ : add_learnt_triplets per graph            base_dataset.py:89-139
    collate             : pad objs / boxes / triplets               packed_coco.py:385-478
    Sg2LayoutModel      : embeddings, 5x GraphTripleConv, box_net   model.py:90-124
    layout              : the generator's own AttributeEmbeddings of the objects, per-image boxes_to_layout on
                          the GT boxes with the dummy objects removed  generator.py:16,80-96, layout.py:12-45
    loss + backward + Adam                                          pix2pix_model.py:72-85, train.py:366-368
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import canon as ocanon, graph as ograph, layout as olayout


def collate(vocab, graphs, canon_out):
    """packed_coco.py:385-478: zero-padded objs, -1 boxes, [0, __padding__, 0] triplets of type 0."""
    B, A = len(graphs), graphs[0].objs.shape[1]
    Omax = max(len(g.objs) for g in graphs)
    Tmax = max(len(t) for t, _ in canon_out)
    objs = np.zeros((B, Omax, A), np.int64)
    boxes = -np.ones((B, Omax, 4), np.float32)
    trips = np.zeros((B, Tmax, 3), np.int64)
    trips[:, :, 1] = vocab.padding_id
    types = np.zeros((B, Tmax), np.int64)
    for b, (g, (tr, ty)) in enumerate(zip(graphs, canon_out)):
        objs[b, :len(g.objs)], boxes[b, :len(g.boxes)] = g.objs, g.boxes
        trips[b, :len(tr)], types[b, :len(ty)] = tr, ty
    return (torch.from_numpy(objs), torch.from_numpy(boxes), torch.from_numpy(trips), torch.from_numpy(types))


def bbox_pred_loss(boxes_pred, boxes, objs, weight=10.0):
    """pix2pix_model.py:72-85: (G_losses["bbox_pred"], G_losses["bbox_pred_all"])."""
    flat = F.smooth_l1_loss(boxes_pred.view(-1, 4), boxes.view(-1, 4), reduction='none') * weight
    fo = objs.view(-1, objs.size(-1))
    mask = (fo.sum(1, keepdim=True) != 0) if objs.size(-1) > 1 else (fo != 0)
    real = mask.to(boxes_pred.dtype)
    per_image = (flat * real).view(boxes.shape).sum(dim=[1, 2]) / real.view(boxes.shape[0], boxes.shape[1]).sum(dim=1)
    return per_image.mean(), per_image


class CpuStep:
    def __init__(self, vocab, state, conv_weights, layout_state, H=64, W=64, learned_converse=True,
                 learned_transitivity=True, lr=1e-4, bbox_pred_loss_weight=10.0):
        self.vocab, self.H, self.W = vocab, H, W
        self.flags = (learned_converse, learned_transitivity)
        self.conv_weights = conv_weights
        self.weight = bbox_pred_loss_weight
        self.state = {k: torch.from_numpy(np.array(v)).requires_grad_(k != "converse_candidates_weights")
                      for k, v in state.items()}
        for k, v in layout_state.items():        # generator.py:16
            self.state["generator.attribute_embedding." + k] = torch.from_numpy(np.array(v)).requires_grad_(True)
        self.opt = torch.optim.Adam([p for p in self.state.values() if p.requires_grad], lr=lr)

    def canonicalize(self, graphs, uniforms, tri_off):
        out = []
        for i, g in enumerate(graphs):
            tr, _, ty, _ = ocanon.add_learnt_triplets(g.triplets, self.vocab.num_preds, self.vocab.meta_ids,
                                                      self.conv_weights, self.flags[0], self.flags[1],
                                                      uniforms[tri_off[i]:])
            out.append((tr, ty))
        return out

    def step(self, graphs, uniforms, canvas_grad=None):
        tri_off = np.concatenate([[0], np.cumsum([len(g.triplets) for g in graphs])])
        canon_out = self.canonicalize(graphs, uniforms, tri_off)
        objs, boxes, trips, types = collate(self.vocab, graphs, canon_out)
        obj_vecs, boxes_pred = ograph.sg2layout_forward(self.state, objs, trips, types, self.vocab.padding_id)
        layout_vecs = ograph.attribute_embeddings(self.state, "generator.attribute_embedding.", objs)   # generator.py:80
        canv = []
        for b in range(len(graphs)):                                  # generator.py:81-96
            keep = (objs[b] != 0)[:, 0]                               # utils.py:56-63 (padding == __image__ == 0)
            canv.append(olayout.boxes_to_layout(layout_vecs[b][keep], boxes[b][keep], self.H, self.W))
        canvas = torch.cat(canv, 0)
        loss, _ = bbox_pred_loss(boxes_pred, boxes, objs, self.weight)
        if canvas_grad is None:
            canvas_grad = torch.ones_like(canvas)
        self.opt.zero_grad(set_to_none=True)
        torch.autograd.backward([canvas, loss], [canvas_grad, None])
        self.opt.step()
        return loss.detach(), sum(len(t) for t, _ in canon_out)
