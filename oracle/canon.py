"""Oracle (numpy) for the WSGC canonicalization step.  TEST INFRASTRUCTURE ONLY.

Follows ``scripts/graphs_utils.py`` and ``sg2im/data/base_dataset.py`` of the
reference.  All matrices are boolean numpy arrays, all triplets int64 arrays
of shape [T, 3] = (subject, predicate, object).

The only randomness of the reference path is ``np.random.choice`` in
``get_edge_converse_triplets`` (graphs_utils.py:143).  Legacy
``RandomState.choice(a, p=p)`` draws exactly one ``random_sample()`` double
``u`` and returns ``a[searchsorted(cumsum(p)/cumsum(p)[-1], u, 'right')]``; the
oracle therefore takes the uniforms as an explicit input, so that results are
defined bit-exactly "given the draws" (``make_golden.py`` checks that feeding
``RandomState(seed).random_sample(n)`` reproduces the reference under
``np.random.seed(seed)``).
"""
import numpy as np

ORIGINAL_EDGE = 0      # base_dataset.py:7
TRANSITIVE_EDGE = 1    # base_dataset.py:8

META_RELATIONS = ("__padding__", "__in_image__")                      # base_dataset.py:14
AUGMENTED_RELATIONS = ("__below__", "__above__", "__left of__",       # base_dataset.py:15
                       "__right of__", "__inside__", "__surrounding__")


# ----------------------------------------------------------------------------
# adjacency helpers
# ----------------------------------------------------------------------------
def triplets_to_adj(triplets):
    """graphs_utils.py:47-55 — N = max index + 1, adj[s, o] = 1."""
    t = np.asarray(triplets, dtype=np.int64)
    n = int(max(t[:, 0].max(), t[:, 2].max())) + 1
    adj = np.zeros((n, n), dtype=bool)
    adj[t[:, 0], t[:, 2]] = True
    return adj


def adj_to_triplets(adj, rel):
    """graphs_utils.py:58-61 — row-major enumeration of set cells.
    (The reference returns float64; the oracle returns int64.)"""
    rows, cols = np.nonzero(np.asarray(adj, dtype=bool))
    out = np.empty((len(rows), 3), dtype=np.int64)
    out[:, 0], out[:, 1], out[:, 2] = rows, int(rel), cols
    return out


def closure(adj):
    """graphs_utils.py:15-27 (``path``) — Warshall in row-OR form.  For pivot
    i every row j != i with adj[j, i] set absorbs row i.  Row i is not written
    during its own pivot step, so the j loop can be taken at once."""
    p = np.array(adj, dtype=bool, copy=True)
    for i in range(p.shape[0]):
        absorb = p[:, i].copy()
        absorb[i] = False
        p[absorb] |= p[i]
    return p


def hsu_reduce(m):
    """graphs_utils.py:30-38 (``hsu``) — sequential, in place, order dependent:
    j outer, i inner, ``row[i] &= ~row[j]`` with the *current* row[j]
    (for i == j this clears row j)."""
    n = m.shape[0]
    for j in range(n):
        for i in range(n):
            if m[i, j]:
                m[i] &= ~m[j]
    return m


def minimal_graph(adj):
    """graphs_utils.py:41-44."""
    return hsu_reduce(closure(adj))


def triplets_to_minimal(triplets):
    """graphs_utils.py:64-71 — identity when fewer than 3 triplets."""
    if len(triplets) < 3:
        return np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    t = np.asarray(triplets, dtype=np.int64)
    return adj_to_triplets(minimal_graph(triplets_to_adj(t)), t[0, 1])


def current_and_transitive(triplets):
    """graphs_utils.py:101-105 — (current edges, closure minus current)."""
    t = np.asarray(triplets, dtype=np.int64)
    cur = triplets_to_adj(t)
    full = closure(cur)
    return adj_to_triplets(cur, t[0, 1]), adj_to_triplets(full & ~cur, t[0, 1])


def minimal_and_transitive(triplets):
    """graphs_utils.py:93-98."""
    t = np.asarray(triplets, dtype=np.int64)
    adj = triplets_to_adj(t)
    mini = minimal_graph(adj)
    full = closure(adj)
    return adj_to_triplets(mini, t[0, 1]), adj_to_triplets(full & ~mini, t[0, 1])


# ----------------------------------------------------------------------------
# converse sampling
# ----------------------------------------------------------------------------
def converse_cdf(conv_weights, rel, candidates):
    """graphs_utils.py:132-139 + legacy ``RandomState.choice`` —
    softmax over [W[rel, c] for c in candidates] + [0], then the normalised
    cumulative sum that ``choice`` searches."""
    logits = np.array([conv_weights[rel, c] for c in candidates] + [0.0], dtype=np.float64)
    e = np.exp(logits - logits.max())          # scipy.special.softmax
    p = e / e.sum()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def converse_table(conv_weights, num_rel, meta_ids):
    """All per-relation CDFs at once: table[rel] is a [num_rel + 1] vector in
    *relation-id space* (cdf value of the last candidate <= id), so that
    ``searchsorted`` over the compact candidate list can be replayed from it.
    Returned as (dist_vals[rel] list, cdf[rel] array) pairs."""
    non_meta = [r for r in range(num_rel) if r not in meta_ids]
    out = {}
    for rel in non_meta:
        cands = [c for c in non_meta if c != rel]
        out[rel] = (cands + [num_rel], converse_cdf(conv_weights, rel, cands))
    return out


def edge_converse_triplets(rel_triplets, candidates, conv_weights, counts, uniforms, cursor):
    """graphs_utils.py:130-155 with the draws made explicit.
    Returns (converse edges, new cursor); ``counts`` is updated in place."""
    rel = int(rel_triplets[0, 1])
    vals = list(candidates) + [counts.shape[1] - 1]
    cdf = converse_cdf(conv_weights, rel, candidates)
    out = []
    for t in rel_triplets:
        u = uniforms[cursor]
        cursor += 1
        r = vals[int(np.searchsorted(cdf, u, side="right"))]
        counts[rel, r] += 1
        if r == counts.shape[1] - 1:
            continue
        out.append(np.array([t[2], r, t[0]], dtype=np.int64))
    return out, cursor


# ----------------------------------------------------------------------------
# add_learnt_triplets
# ----------------------------------------------------------------------------
def add_learnt_triplets(triplets, num_rel, meta_ids, conv_weights=None,
                        learned_converse=False, learned_transitivity=False,
                        uniforms=None):
    """base_dataset.py:89-139.

    Returns (triplets' [T',3] int64, conv_counts [P, P+1] float64,
             triplet_type [T'] int64, number of uniforms consumed)."""
    trip = np.unique(np.asarray(triplets), axis=0).astype(np.int64)          # :90
    counts = np.zeros((num_rel, num_rel + 1))                                # :93
    meta_ids = list(meta_ids)
    non_meta = [r for r in range(num_rel) if r not in meta_ids]              # :96-97 (ascending)
    cursor = 0
    new = []
    for rel in non_meta:                                                     # :98-110
        rel_t = trip[trip[:, 1] == rel]
        if len(rel_t) == 0:
            continue
        new.extend(rel_t)
        if learned_converse:
            conv, cursor = edge_converse_triplets(
                rel_t, [c for c in non_meta if c != rel], conv_weights, counts, uniforms, cursor)
            new.extend(conv)
    transitive = []
    if learned_transitivity:                                                 # :112-122
        arr = np.array(new, dtype=np.int64).reshape(-1, 3)
        for rel in non_meta:
            if not len(arr):
                continue
            rel_t = arr[arr[:, 1] == rel]
            if not len(rel_t):
                continue
            transitive.extend(current_and_transitive(rel_t)[1])
    new = list(new)
    for rel in meta_ids:                                                     # :124-127
        new.extend(trip[trip[:, 1] == rel])
    out = np.unique(np.array(new, dtype=np.int64).reshape(-1, 3), axis=0)     # :129-130
    types = [ORIGINAL_EDGE] * len(out)
    if len(transitive) > 0:                                                  # :134-137
        types = types + [TRANSITIVE_EDGE] * len(transitive)
        out = np.concatenate([out, np.array(transitive, dtype=np.int64)], axis=0)
    return out, counts, np.array(types, dtype=np.int64), cursor


# ----------------------------------------------------------------------------
# add_location_triplets / add_dummy_triplets
# ----------------------------------------------------------------------------
def add_location_triplets(boxes, obj_centers, objs, image_obj_id, pred_ids):
    """base_dataset.py:35-87.  ``boxes`` [O,4] xywh float32, ``obj_centers``
    [O,2] float32, ``objs`` [O] int, ``pred_ids`` maps the six augmented
    relation names to ids.  Returns the list of minimal per-relation triplets
    in the reference's emission order."""
    boxes = np.asarray(boxes, dtype=np.float32)
    cen = np.asarray(obj_centers, dtype=np.float32)
    objs = np.asarray(objs)
    real = [int(i) for i in np.nonzero(objs != image_obj_id)[0]] if len(objs) > 1 else []
    two = np.float32(2)
    raw = []
    for s in real:
        for o in real:
            if o == s:
                continue
            sx0, sy0, sw, sh = boxes[s]
            sx1, sy1 = sx0 + sw / two, sy0 + sh / two                        # :47 (sic: x0 + w/2)
            ox0, oy0, ow, oh = boxes[o]
            ox1, oy1 = ox0 + ow / two, oy0 + oh / two
            d = cen[s] - cen[o]
            if sx0 < ox0 and sx1 > ox1 and sy0 < oy0 and sy1 > oy1:
                raw.append([s, pred_ids["__surrounding__"], o])
            elif sx0 > ox0 and sx1 < ox1 and sy0 > oy0 and sy1 < oy1:
                raw.append([s, pred_ids["__inside__"], o])
            else:
                if d[0] > 0:
                    raw.append([s, pred_ids["__right of__"], o])
                elif d[0] < 0:
                    raw.append([s, pred_ids["__left of__"], o])
                if d[1] > 0:
                    raw.append([s, pred_ids["__below__"], o])
                elif d[1] < 0:
                    raw.append([s, pred_ids["__above__"], o])
    raw = np.array(raw, dtype=np.int64).reshape(-1, 3)
    out = []
    for name in AUGMENTED_RELATIONS:                                         # :82-87
        p = pred_ids[name]
        out.extend(triplets_to_minimal(raw[raw[:, 1] == p]))
    return [np.asarray(t, dtype=np.int64) for t in out]


def add_dummy_triplets(objs, image_obj_id, in_image_id, include_dummies=True):
    """base_dataset.py:141-150 — [i, __in_image__, image] for every other object."""
    out = []
    if include_dummies:
        objs = np.asarray(objs)
        img = int(np.nonzero(objs == image_obj_id)[0][0])
        for i in range(len(objs)):
            if i != img:
                out.append(np.array([i, in_image_id, img], dtype=np.int64))
    return out


def make_vocab_pred_ids(base_predicates=()):
    """base_dataset.py:152-161 — ids: dataset predicates first (VG), then the
    two meta relations, then the six augmented relations."""
    names = list(base_predicates)
    for p in META_RELATIONS + AUGMENTED_RELATIONS:
        if p not in names:
            names.append(p)
    return {n: i for i, n in enumerate(names)}, names
