"""CPU restatement of the reference's padded collate functions (TEST INFRASTRUCTURE ONLY).

``sg2im/data/packed_coco.py:385-478`` (``coco_collate_fn``) and ``sg2im/data/packed_vg.py:147-229``
(``vg_collate_fn``) are the same code up to the masks: every sample is padded to the largest object / triplet count of
the batch (objects with class 0 and box -1, masks with zeros, triplets with ``[0, __padding__, 0]`` of type 0) and the
padded samples are stacked.  Pinned by ``oracle/make_golden.py`` (``tests/golden/collate.npz``: outputs of the
unmodified reference functions on seeded samples).
"""
import torch


def padded_collate(vocab, batch, with_masks=True):
    """batch: list of ``(img, objs: dict attr -> LongTensor[O], boxes [O,4], triplets [T,3], conv_counts, triplet_type [T],
    masks [O,M,M] or None, image_id)``.  Returns the reference's 8-tuple."""
    max_o = max(next(iter(s[1].values())).size(0) for s in batch)              # packed_coco.py:408-416
    max_t = max(s[3].size(0) for s in batch)
    pad_pred = vocab["pred_name_to_idx"]["__padding__"]
    imgs, all_objs, all_boxes, all_trip, all_cc, all_ty, all_masks, ids = [], [], [], [], [], [], [], []
    for img, objs, boxes, triplets, conv_counts, triplet_type, masks, image_id in batch:
        O, T = next(iter(objs.values())).size(0), triplets.size(0)
        attrs = list(objs.keys())                                               # :424-426 (the sort is discarded)
        ao = torch.zeros(len(attrs), max_o, dtype=torch.long)
        for k, v in objs.items():
            ao[attrs.index(k), :O] = v                                          # :429-436
        all_objs.append(ao.transpose(1, 0))
        if max_o - O > 0:                                                       # :439-441
            boxes = torch.cat([boxes, torch.full((max_o - O, 4), -1.0)])
            if with_masks and masks is not None:                                # :444-446
                masks = torch.cat([masks, torch.zeros((max_o - O, masks.size(1), masks.size(2)), dtype=torch.long)])
        if max_t - T > 0:                                                       # :449-453
            triplets = torch.cat([triplets, torch.tensor([[0, pad_pred, 0]], dtype=torch.long).repeat(max_t - T, 1)])
            triplet_type = torch.cat([triplet_type, torch.zeros(max_t - T, dtype=torch.long)])
        imgs.append(img[None]); ids.append(image_id)
        all_boxes.append(boxes); all_trip.append(triplets); all_ty.append(triplet_type); all_cc.append(conv_counts)
        if with_masks and masks is not None and all_masks is not None:
            all_masks.append(masks)
        else:
            all_masks = None                                                    # :458-461
    return (torch.cat(imgs), torch.stack(all_objs), torch.stack(all_boxes), torch.stack(all_trip),
            torch.stack(all_cc).to(torch.float32), torch.stack(all_ty),
            torch.stack(all_masks) if all_masks is not None else None, torch.LongTensor(ids))
