"""canonicalsg2im_b200 -- B200 (sm_100a) kernels for the scene-graph -> layout hot path of
roeiherz/CanonicalSg2Im, behind the reference's own Python operator surface.

    from canonicalsg2im_b200 import GraphTripleConv, GraphTripleConvNet, Sg2LayoutModel
    from canonicalsg2im_b200 import boxes_to_layout, masks_to_layout, layout_batched
    from canonicalsg2im_b200 import add_learnt_triplets, add_learnt_triplets_batched, add_location_triplets_batched
    from canonicalsg2im_b200 import crop_bbox, crop_bbox_batch

Everything computes in ``libcsg2im.so`` (hand-written CUDA, C ABI in ``include/csg2im.h``); importing
the operators without the built library, or calling them on CPU tensors, raises.
"""
from . import synth  # noqa: F401  (host-side synthetic inputs; no kernels)

__all__ = ["synth", "GraphTripleConv", "GraphTripleConvNet", "TripleBatch", "Sg2LayoutModel", "get_conv_converse",
           "boxes_to_layout", "masks_to_layout", "layout_batched", "layout_pyramid", "add_learnt_triplets",
           "add_learnt_triplets_batched", "add_location_triplets_batched", "canon_count_async", "canon_emit",
           "converse_tables", "closure", "crop_bbox", "crop_bbox_batch", "crop_bbox_ragged", "FusedAdam"]

_LAZY = {
    "GraphTripleConv": "graph", "GraphTripleConvNet": "graph", "TripleBatch": "graph",
    "Sg2LayoutModel": "model", "get_conv_converse": "model",
    "boxes_to_layout": "layout", "masks_to_layout": "layout", "layout_batched": "layout", "layout_pyramid": "layout",
    "add_learnt_triplets": "canonicalize", "add_learnt_triplets_batched": "canonicalize",
    "add_location_triplets_batched": "canonicalize", "canon_count_async": "canonicalize", "canon_emit": "canonicalize",
    "converse_tables": "canonicalize", "closure": "canonicalize", "FusedAdam": "optim",
    "crop_bbox": "bilinear", "crop_bbox_batch": "bilinear", "crop_bbox_ragged": "bilinear",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    raise AttributeError(name)
