"""Collate straight to the flat / ragged batch the csg2im kernels consume.

The reference pads every sample to the largest object / triplet count of the batch on the CPU
(``sg2im/data/packed_coco.py:385-478``, ``sg2im/data/packed_vg.py:147-229``) and the model un-pads on the GPU.
``ragged_collate_fn`` takes the same per-sample tuples and emits concatenated arrays plus int32 offsets -- no padding is
ever materialised, and ``Sg2LayoutModel.forward_ragged`` / ``layout_batched`` / ``add_learnt_triplets_batched`` take
the result as is.  ``ragged_to_padded`` reproduces the reference's padded 8-tuple bit for bit (drop-in check, and a
bridge to reference code that still wants padded tensors).  Host-side code: no kernels here.
"""
import torch


class RaggedBatch(dict):
    """Flat batch: ``imgs [B,C,H,W]``, ``objs [sum O, A]`` i64, ``boxes [sum O, 4]`` f32, ``masks [sum O, M, M]`` or
    None, ``triplets [sum T, 3]`` i64 with graph-local object ids, ``triplet_type [sum T]`` i64, ``conv_counts
    [B, ...]`` f32, ``obj_off`` / ``tri_off`` ``[B+1]`` i32, ``image_ids [B]`` i64, and the host ints ``B``,
    ``max_objs``, ``max_triplets`` (sizing hints for the kernels)."""

    TENSORS = ("imgs", "objs", "boxes", "masks", "triplets", "triplet_type", "conv_counts", "obj_off", "tri_off",
               "image_ids")

    def pin(self):
        for k in self.TENSORS:
            if torch.is_tensor(self.get(k)):
                self[k] = self[k].pin_memory()
        return self

    def to(self, device, non_blocking=True):
        out = RaggedBatch(self)
        for k in self.TENSORS:
            if torch.is_tensor(self.get(k)):
                out[k] = self[k].to(device, non_blocking=non_blocking)
        return out


def ragged_collate_fn(vocab, batch):
    """Same input as ``coco_collate_fn`` / ``vg_collate_fn``: a list of ``(img, objs: dict attribute -> LongTensor[O],
    boxes [O,4], triplets [T,3], conv_counts, triplet_type [T], masks [O,M,M] or None, image_id)``."""
    del vocab                                     # kept for signature compatibility: nothing is padded
    imgs, objs, boxes, masks, trips, types, ccs, ids = [], [], [], [], [], [], [], []
    obj_off, tri_off = [0], [0]
    has_masks = True
    for img, o, b, t, cc, ty, m, image_id in batch:
        attrs = list(o.keys())                    # the reference's attribute order (its sort is discarded, :424-426)
        objs.append(torch.stack([o[k].to(torch.long) for k in attrs], dim=1))
        O, T = objs[-1].size(0), t.size(0)
        imgs.append(img[None]); boxes.append(b.to(torch.float32)); trips.append(t.to(torch.long))
        types.append(ty.to(torch.long)); ccs.append(cc); ids.append(int(image_id))
        if m is None:
            has_masks = False
        elif has_masks:
            masks.append(m)
        obj_off.append(obj_off[-1] + O); tri_off.append(tri_off[-1] + T)
    rb = RaggedBatch(
        imgs=torch.cat(imgs), objs=torch.cat(objs), boxes=torch.cat(boxes),
        masks=torch.cat(masks) if has_masks and masks else None,
        triplets=torch.cat(trips).reshape(-1, 3), triplet_type=torch.cat(types),
        conv_counts=torch.stack(ccs).to(torch.float32),
        obj_off=torch.tensor(obj_off, dtype=torch.int32), tri_off=torch.tensor(tri_off, dtype=torch.int32),
        image_ids=torch.tensor(ids, dtype=torch.long))
    rb["B"] = len(batch)
    rb["max_objs"] = max(b - a for a, b in zip(obj_off[:-1], obj_off[1:])) if batch else 0
    rb["max_triplets"] = max(b - a for a, b in zip(tri_off[:-1], tri_off[1:])) if batch else 0
    return rb


def ragged_to_padded(rb, padding_pred_id):
    """The reference's padded 8-tuple ``(imgs, objs [B,O,A], boxes [B,O,4], triplets [B,T,3], conv_counts,
    triplet_type [B,T], masks [B,O,M,M] or None, image_ids)`` from a :class:`RaggedBatch` (host tensors)."""
    B, O, T = rb["B"], rb["max_objs"], rb["max_triplets"]
    oo, to = rb["obj_off"].tolist(), rb["tri_off"].tolist()
    objs = torch.zeros((B, O, rb["objs"].size(1)), dtype=torch.long)
    boxes = torch.full((B, O, 4), -1.0)
    trip = torch.zeros((B, T, 3), dtype=torch.long)
    trip[:, :, 1] = int(padding_pred_id)
    types = torch.zeros((B, T), dtype=torch.long)
    masks = None
    if rb.get("masks") is not None:
        m = rb["masks"]
        masks = torch.zeros((B, O) + tuple(m.shape[1:]), dtype=m.dtype)
    for b in range(B):
        no, nt = oo[b + 1] - oo[b], to[b + 1] - to[b]
        objs[b, :no] = rb["objs"][oo[b]:oo[b + 1]]
        boxes[b, :no] = rb["boxes"][oo[b]:oo[b + 1]]
        trip[b, :nt] = rb["triplets"][to[b]:to[b + 1]]
        types[b, :nt] = rb["triplet_type"][to[b]:to[b + 1]]
        if masks is not None:
            masks[b, :no] = rb["masks"][oo[b]:oo[b + 1]]
    return rb["imgs"], objs, boxes, trip, rb["conv_counts"], types, masks, rb["image_ids"]
