"""WSGC canonicalization on the GPU: batched drop-in for ``BaseDataset.add_learnt_triplets``
(reference ``sg2im/data/base_dataset.py:89-139`` with ``scripts/graphs_utils.py:15-155``).

The reference canonicalises one graph at a time on the CPU inside ``Dataset.__getitem__``; here a
whole batch of graphs is completed by two launches of an integer bitset kernel (csrc/canon.cu).
Results are bit-exact given the converse draws: pass the doubles that ``np.random.choice`` would
have consumed (one per unique non-meta triple, in the reference's visiting order) as ``uniforms``.

Host-side pieces (tiny, weight-dependent only): the per-relation softmax CDF table
(``graphs_utils.py:132-139``), built in float64 with numpy exactly as the reference does.
"""
import numpy as np
import torch

from . import _lib
from .ops import lib, ptr, need_cuda, _stream

ORIGINAL_EDGE, TRANSITIVE_EDGE = 0, 1          # base_dataset.py:7-8


def converse_tables(conv_weights, num_rel, meta_ids):
    """CDF / value tables of the converse sampler.

    For every non-meta relation ``rel`` the candidates are the other non-meta relations (ascending)
    plus the "no edge" outcome ``num_rel`` with logit 0 (``graphs_utils.py:132-137``); probabilities
    are a float64 softmax and ``np.random.choice`` searches their normalised cumulative sum.
    Returns (cdf [P, ncand] float64, vals [P, ncand] int32)."""
    W = np.asarray(conv_weights.detach().cpu() if torch.is_tensor(conv_weights) else conv_weights, dtype=np.float64)
    non_meta = [r for r in range(num_rel) if r not in tuple(meta_ids)]
    ncand = len(non_meta)          # (len(non_meta) - 1) candidates + "no edge"
    cdf = np.ones((num_rel, max(ncand, 1)), dtype=np.float64)
    vals = np.full((num_rel, max(ncand, 1)), num_rel, dtype=np.int32)
    for rel in non_meta:
        cands = [c for c in non_meta if c != rel]
        logits = np.array([W[rel, c] for c in cands] + [0.0], dtype=np.float64)
        e = np.exp(logits - logits.max())
        p = e / e.sum()
        c = p.cumsum()
        c /= c[-1]
        cdf[rel, :len(c)] = c
        vals[rel, :len(cands)] = cands
    return cdf, vals


class CanonResult:
    def __init__(self, triplets, triplet_type, tri_off, conv_counts):
        self.triplets = triplets            # [NT', 3] int64, graph-local object ids
        self.triplet_type = triplet_type    # [NT'] int64 (0 original, 1 transitive)
        self.tri_off = tri_off              # [B+1] int32
        self.conv_counts = conv_counts      # [B, P, P+1] int32

    def split(self):
        """Per-graph (triplets, conv_counts, triplet_type) like the reference returns them."""
        off = self.tri_off.cpu().tolist()
        return [(self.triplets[off[b]:off[b + 1]], self.conv_counts[b], self.triplet_type[off[b]:off[b + 1]])
                for b in range(len(off) - 1)]


class CanonPlan:
    """First half of a batched canonicalization (sizes counted, output offsets scanned, sizes on their way to the
    host): what :func:`canon_count_async` returns and :func:`canon_emit` consumes."""

    def __init__(self, args, keep, B, out_off, conv_counts, summary_host, event, max_objs):
        self.args, self.keep, self.B = args, keep, B
        self.out_off, self.conv_counts = out_off, conv_counts
        self.summary_host, self.event, self.max_objs = summary_host, event, max_objs


def canon_count_async(triplets, tri_off, obj_off, num_rel, meta_ids, conv_weights=None,
                      learned_converse=False, learned_transitivity=False, uniforms=None,
                      max_objs_per_graph=None, tables=None):
    """Counting pass of :func:`add_learnt_triplets_batched` without a host synchronisation: launches
    ``csg_canon_count`` + the offset scan and an asynchronous copy of the two summary integers into pinned host
    memory.  A data pipeline calls this for batch i+1 while batch i trains, so that :func:`canon_emit` finds the
    sizes already on the host and the launch queue never drains."""
    need_cuda(triplets, tri_off, obj_off, uniforms)
    dev = triplets.device
    L = lib()
    B = tri_off.numel() - 1
    tr = triplets.contiguous().to(torch.int64)
    tri_off = tri_off.to(torch.int32).contiguous()
    obj_off = obj_off.to(torch.int32).contiguous()
    meta = list(meta_ids) + [-1, -1]
    if max_objs_per_graph is None:
        max_objs_per_graph = int((obj_off[1:] - obj_off[:-1]).max().item()) if B else 1
    cdf_t = vals_t = None
    ncand = 0
    if learned_converse:
        if tables is None:
            tables = converse_tables(conv_weights, num_rel, meta_ids)
        cdf, vals = tables
        cdf_t = cdf if torch.is_tensor(cdf) else torch.from_numpy(cdf)
        vals_t = vals if torch.is_tensor(vals) else torch.from_numpy(vals)
        cdf_t, vals_t = cdf_t.to(dev).contiguous(), vals_t.to(dev).contiguous()
        ncand = cdf_t.shape[1]
        if uniforms is None or uniforms.numel() < tr.shape[0]:
            raise ValueError("learned_converse needs one float64 draw per input triple")
        uniforms = uniforms.to(torch.float64).contiguous()
    cnt = torch.empty((2, max(B, 1)), dtype=torch.int32, device=dev)
    conv_counts = torch.empty((max(B, 1), num_rel, num_rel + 1), dtype=torch.int32, device=dev)
    args = (ptr(tr), ptr(tri_off), ptr(obj_off), B, ptr(uniforms) if learned_converse else 0, ptr(cdf_t), ptr(vals_t),
            ncand, num_rel, meta[0], meta[1], int(learned_converse), int(learned_transitivity), int(max_objs_per_graph))
    _lib.check(L.csg_canon_count(*args, ptr(cnt[0]), ptr(cnt[1]), ptr(conv_counts), _stream()), "csg_canon_count")
    out_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    summary = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(L.csg_canon_offsets(ptr(cnt[0]), ptr(cnt[1]), B, ptr(out_off), ptr(summary), _stream()),
               "csg_canon_offsets")
    summary_host = torch.empty(2, dtype=torch.int32, pin_memory=True)
    summary_host.copy_(summary, non_blocking=True)
    event = torch.cuda.Event()
    event.record()
    keep = (tr, tri_off, obj_off, uniforms, cdf_t, vals_t, cnt, summary)     # device buffers the emit pass reads
    return CanonPlan(args, keep, B, out_off, conv_counts, summary_host, event, max_objs_per_graph)


def canon_total(plan):
    """Number of output triples of ``plan`` (waits on the host for the counting pass)."""
    plan.event.synchronize()                   # sizes the output allocation (the one host wait per batch)
    total, min_cnt0 = plan.summary_host.tolist()
    if plan.B and min_cnt0 < 0:
        raise _lib.CsgError("canonicalize: a graph has more objects than max_objs_per_graph=%d" % plan.max_objs)
    return total


def canon_emit(plan, out=None):
    """Second half: waits (on the host) for the sizes of ``plan`` -- already there when the plan was launched a step
    ahead --, allocates the output (or uses ``out = (triplets [>= total, 3], types [>= total])`` int64 buffers) and
    launches ``csg_canon_emit``.  Returns a :class:`CanonResult`."""
    total = canon_total(plan)
    dev = plan.out_off.device
    # the counting pass may have run on another stream (a data pipeline's look-ahead stream): its buffers were allocated
    # there, so tell the caching allocator that this stream reads them now (canon_total has already waited for the pass)
    cur = torch.cuda.current_stream()
    for x in plan.keep + (plan.out_off, plan.conv_counts):
        if torch.is_tensor(x) and x.is_cuda:
            x.record_stream(cur)
    if out is not None:
        out_t, out_ty = out
        if out_t.shape[0] < total or out_ty.shape[0] < total or out_t.dtype != torch.int64 or not out_t.is_contiguous():
            raise ValueError("canon_emit: `out` buffers must be contiguous int64 with at least %d rows" % total)
    else:
        out_t = torch.empty((max(total, 1), 3), dtype=torch.int64, device=dev)
        out_ty = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    _lib.check(lib().csg_canon_emit(*plan.args, ptr(plan.out_off), ptr(out_t), ptr(out_ty), _stream()), "csg_canon_emit")
    return CanonResult(out_t[:total], out_ty[:total], plan.out_off, plan.conv_counts[:plan.B])


def add_learnt_triplets_batched(triplets, tri_off, obj_off, num_rel, meta_ids, conv_weights=None,
                                learned_converse=False, learned_transitivity=False, uniforms=None,
                                max_objs_per_graph=None, tables=None):
    """Canonicalise B graphs at once.

    triplets [NTin, 3] int64 (graph-local ids, any order, duplicates allowed), tri_off / obj_off [B+1]
    int32 (CUDA), ``uniforms`` [>= NTin] float64 (CUDA): the k-th converse draw of graph g is
    ``uniforms[tri_off[g] + k]``.  Returns a :class:`CanonResult` (device tensors)."""
    return canon_emit(canon_count_async(triplets, tri_off, obj_off, num_rel, meta_ids, conv_weights, learned_converse,
                                        learned_transitivity, uniforms, max_objs_per_graph, tables))


AUGMENTED_RELATIONS = ("__below__", "__above__", "__left of__", "__right of__", "__inside__", "__surrounding__")


def add_location_triplets_batched(boxes, obj_centers, objs, obj_off, image_obj_id, pred_ids, max_objs_per_graph=None,
                                  in_image_pred=None):
    """``BaseDataset.add_location_triplets`` (base_dataset.py:35-87) for a whole flat batch on the device; with
    ``in_image_pred`` (the id of ``__in_image__``) also ``add_dummy_triplets`` (base_dataset.py:141-150): every graph's
    rows are its location triplets followed by ``[i, __in_image__, image]`` for its objects, i.e. the ``triplets`` list
    the reference datasets hand to ``add_learnt_triplets`` (packed_coco.py:355-357).

    boxes [NO, 4] xywh float32, obj_centers [NO, 2] float32, objs [NO] or [NO, A] int64 class ids (column 0 is used),
    obj_off [B+1] int32 (all CUDA); ``pred_ids`` maps the six augmented relation names to predicate ids.  Returns
    ``(triplets [T, 3] int64 with graph-local ids, tri_off [B+1] int32)``: per graph, the rows the reference appends to
    ``triplets`` (relation by relation in ``augmented_relations`` order, each reduced by ``triplets_to_minimal``)."""
    need_cuda(boxes, obj_centers, objs, obj_off)
    dev = boxes.device
    L = lib()
    B = obj_off.numel() - 1
    bx = boxes.to(torch.float32).contiguous()
    cen = obj_centers.to(torch.float32).contiguous()
    ob = objs.to(torch.int64)
    if ob.dim() == 2:
        ob = ob[:, 0]
    off = obj_off.to(torch.int32).contiguous()
    if max_objs_per_graph is None:
        max_objs_per_graph = int((off[1:] - off[:-1]).max().item()) if B else 1
    import ctypes
    pid = (ctypes.c_int * 6)(*[int(pred_ids[name]) for name in AUGMENTED_RELATIONS])
    cnt = torch.zeros((2, max(B, 1)), dtype=torch.int32, device=dev)
    args = (ptr(bx), ptr(cen), ptr(ob), ob.stride(0) if ob.numel() else 1, ptr(off), B, int(image_obj_id), pid,
            int(max(max_objs_per_graph, 1)))
    _lib.check(L.csg_location_count(*args, ptr(cnt[0]), _stream()), "csg_location_count")
    if in_image_pred is not None:
        _lib.check(L.csg_dummy_triplets_count(ptr(ob), ob.stride(0) if ob.numel() else 1, ptr(off), B, int(image_obj_id),
                                              ptr(cnt[1]), _stream()), "csg_dummy_triplets_count")
    out_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    summary = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(L.csg_canon_offsets(ptr(cnt[0]), ptr(cnt[1]), B, ptr(out_off), ptr(summary), _stream()), "csg_canon_offsets")
    total, min_cnt = summary.tolist()
    if B and min_cnt < 0:
        raise _lib.CsgError("location triplets: a graph has more objects than max_objs_per_graph=%d" % max_objs_per_graph)
    out = torch.empty((max(total, 1), 3), dtype=torch.int64, device=dev)
    _lib.check(L.csg_location_emit(*args, ptr(out_off), ptr(out), _stream()), "csg_location_emit")
    if in_image_pred is not None:
        _lib.check(L.csg_dummy_triplets_emit(ptr(ob), ob.stride(0) if ob.numel() else 1, ptr(off), B, int(image_obj_id),
                                             int(in_image_pred), ptr(out_off), ptr(cnt[0]), ptr(out), _stream()),
                   "csg_dummy_triplets_emit")
    return out[:total], out_off


def add_learnt_triplets(triplets, O, num_rel, meta_ids, conv_weights=None, learned_converse=False,
                        learned_transitivity=False, uniforms=None, device="cuda"):
    """Single-graph form mirroring ``BaseDataset.add_learnt_triplets(triplets, O)`` (base_dataset.py:89):
    returns (triplets' [T',3] int64 ndarray, conv_counts [P,P+1] float64 ndarray, triplet_type list)."""
    t = torch.as_tensor(np.asarray(triplets, dtype=np.int64).reshape(-1, 3), device=device)
    n = t.shape[0]
    tri_off = torch.tensor([0, n], dtype=torch.int32, device=device)
    obj_off = torch.tensor([0, int(O)], dtype=torch.int32, device=device)
    u = None
    if learned_converse:
        u = torch.zeros(max(n, 1), dtype=torch.float64, device=device)
        src = torch.as_tensor(np.asarray(uniforms, dtype=np.float64), device=device)
        k = min(n, src.numel())
        u[:k] = src[:k]
    res = add_learnt_triplets_batched(t, tri_off, obj_off, num_rel, meta_ids, conv_weights, learned_converse,
                                      learned_transitivity, u, max_objs_per_graph=int(O))
    return (res.triplets.cpu().numpy(), res.conv_counts[0].cpu().numpy().astype(np.float64),
            res.triplet_type.cpu().tolist())


def closure(adj, reduce=False):
    """``path`` (graphs_utils.py:15-27) or, with ``reduce``, ``get_minimal_graph`` (:41-44) of a batch
    of adjacency matrices [G, n, n] (uint8/bool CUDA tensor)."""
    need_cuda(adj)
    a = adj.to(torch.uint8).contiguous()
    squeeze = a.dim() == 2
    if squeeze:
        a = a[None]
    out = torch.empty_like(a)
    _lib.check(lib().csg_canon_closure(ptr(a), a.shape[0], a.shape[1], int(reduce), ptr(out), _stream()),
               "csg_canon_closure")
    return out[0] if squeeze else out
