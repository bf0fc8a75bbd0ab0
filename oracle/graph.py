"""Oracle (torch CPU) for GraphTripleConv and the Sg2LayoutModel GCN stack.
TEST INFRASTRUCTURE ONLY.

Restates ``sg2im/graph.py:44-113`` and ``sg2im/model.py:90-124`` of the
reference op for op (same per-sample Python loops, same scatter order), as
pure functions over a ``state`` dict that uses the reference's state-dict
keys (``net1.0.weight`` ...).  Autograd works through it, so it is also the
fwd+bwd CPU baseline ("port") that ``bench.py`` times.
"""
import torch
import torch.nn.functional as F

ORIGINAL_EDGE = 0      # sg2im/data/base_dataset.py:7
TRANSITIVE_EDGE = 1    # sg2im/data/base_dataset.py:8


def mlp2(x, w0, b0, w1, b1, final_relu=True):
    """sg2im/layers.py:6-25 with dim_list of length 3, batch_norm='none':
    Linear, ReLU, Linear, [ReLU]."""
    y = F.linear(F.relu(F.linear(x, w0, b0)), w1, b1)
    return F.relu(y) if final_relu else y


def triple_confidence(triplet_type, predicate_ids, w_trans, dtype):
    """sg2im/graph.py:69-74."""
    tt = triplet_type.to(dtype)
    sig = torch.sigmoid(w_trans)
    return (tt == ORIGINAL_EDGE).to(dtype) + (tt == TRANSITIVE_EDGE).to(dtype) * sig[predicate_ids]


def graph_triple_conv(state, prefix, obj_vecs, pred_vecs, edges, pred_indicators,
                      triplet_type, predicate_ids, w_trans, hidden_dim, pred_out_dim,
                      return_new_p_vecs=True):
    """sg2im/graph.py:44-113.

    obj_vecs [B,O,Din], pred_vecs [B,T,Dp], edges [B,T,2] i64,
    pred_indicators [B,T] bool, triplet_type [B,T] i64, predicate_ids [B,T] i64.
    Returns (new_obj_vecs [B,O,Dout], new_p_vecs [B,T,Dp_out])."""
    g = lambda k: state[prefix + k]
    dtype = obj_vecs.dtype
    B, O = obj_vecs.shape[0], obj_vecs.shape[1]
    H = hidden_dim
    s_idx = edges[:, :, 0].contiguous()
    o_idx = edges[:, :, 1].contiguous()
    cur_s = torch.stack([obj_vecs[b, s_idx[b]] for b in range(B)])           # :63
    cur_o = torch.stack([obj_vecs[b, o_idx[b]] for b in range(B)])           # :64
    t_in = torch.cat([cur_s, pred_vecs, cur_o], dim=-1)                      # :66
    t_out = mlp2(t_in, g("net1.0.weight"), g("net1.0.bias"),
                 g("net1.2.weight"), g("net1.2.bias"))                       # :67
    conf = triple_confidence(triplet_type, predicate_ids, w_trans, dtype)    # :69-74
    t_out = t_out * conf.unsqueeze(-1)                                       # :76-77
    new_s = t_out[:, :, :H]                                                  # :79-81
    new_p = t_out[:, :, H:H + pred_out_dim]
    new_o = t_out[:, :, H + pred_out_dim:]
    pooled_all = []
    for b in range(B):                                                       # :85-107
        keep = pred_indicators[b]
        si, oi = s_idx[b][keep], o_idx[b][keep]
        vs, vo, cf = new_s[b][keep], new_o[b][keep], conf[b][keep]
        pooled = torch.zeros(O, H, dtype=dtype)
        pooled = pooled.scatter_add(0, si.view(-1, 1).expand_as(vs), vs)
        pooled = pooled.scatter_add(0, oi.view(-1, 1).expand_as(vo), vo)
        cnt = torch.zeros(O, dtype=dtype)
        cnt = cnt.scatter_add(0, si, cf)
        cnt = cnt.scatter_add(0, oi, cf)
        nz = cnt > 0
        pooled[nz] = pooled[nz] / cnt[nz].view(-1, 1)                        # :105-106 (masked in-place divide)
        pooled_all.append(pooled)
    pooled_all = torch.stack(pooled_all, dim=0)
    new_obj = mlp2(pooled_all, g("net2.0.weight"), g("net2.0.bias"),
                   g("net2.2.weight"), g("net2.2.bias"))                     # :109-110
    if not return_new_p_vecs:
        new_p = pred_vecs
    return new_obj, new_p


def attribute_embeddings(state, prefix, objs):
    """sg2im/attribute_embed.py:32-48 — per-attribute lookup, concat, optional FC."""
    vecs = [F.embedding(objs[:, :, k], state[prefix + "att_emb_%d.weight" % k])
            for k in range(objs.shape[-1])]
    v = torch.cat(vecs, dim=-1)
    if prefix + "attribute_fc_gen.weight" in state:
        v = F.linear(v, state[prefix + "attribute_fc_gen.weight"], state[prefix + "attribute_fc_gen.bias"])
    return v


def sg2layout_forward(state, objs, triplets, triplet_type, padding_pred_id,
                      num_layers=5, hidden_dim=512, gconv_dim=128):
    """sg2im/model.py:90-124 without the optional mask_net.

    objs [B,O,A] i64, triplets [B,T,3] i64, triplet_type [B,T] i64.
    Returns (obj_vecs [B,O,gconv_dim], boxes_pred [B,O,4])."""
    s, p, o = triplets[:, :, 0], triplets[:, :, 1], triplets[:, :, 2]        # :104-105
    edges = torch.stack([s, o], dim=-1)                                      # :106
    pred_indicators = p != padding_pred_id                                   # :107
    obj_vecs = attribute_embeddings(state, "attribute_embedding.", objs)     # :108
    pred_vecs = F.embedding(p, state["pred_embeddings.weight"])              # :109
    w_trans = state["trans_candidates_weights"]
    for i in range(num_layers):                                              # :111-112
        obj_vecs, pred_vecs = graph_triple_conv(
            state, "gconvs.%d." % i, obj_vecs, pred_vecs, edges, pred_indicators,
            triplet_type, p, w_trans, hidden_dim, gconv_dim)
    boxes = mlp2(obj_vecs, state["box_net.0.weight"], state["box_net.0.bias"],
                 state["box_net.2.weight"], state["box_net.2.bias"], final_relu=False)  # :115
    return obj_vecs, boxes
