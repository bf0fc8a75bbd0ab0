"""Plain-PyTorch fp32 reference of the tensor-core engine's arithmetic: the reference layer
(sg2im/graph.py:44-113) with operands rounded to bf16 at the points where the kernels store bf16
(inputs, weights, hidden, net1 output, pooled, net2 hidden, output) and fp32 accumulation in between.
Rounding is straight-through for autograd, so its gradients have the same ReLU masks as the kernels'
and differ only by the bf16 rounding of intermediate gradients."""
import torch


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def r(x):
    return _Round.apply(x)


class _RoundGrad(torch.autograd.Function):
    """identity forward; the incoming gradient is rounded to bf16 -- the points where the backward kernels store a
    gradient tensor in bf16 (g4, dh2, g, dhid, dX, dobj of the inner layers; csrc/gconv_engine.cu)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def rg(x, on=True):
    return _RoundGrad.apply(x) if on else x


def layer_bf16_ref(state, obj, pred, s_idx, o_idx, valid, conf_fn, H, Dpo, round_grads=False, obj_grad_bf16=False):
    """Flat tensors: obj [NO, Din], pred [NT, Dp], s_idx / o_idx [NT] long (global), valid [NT] bool.
    ``round_grads``: also round the gradient tensors the backward kernels store in bf16 (``obj_grad_bf16``: the
    gradient wrt the object rows too, as for the layers whose input is the previous layer's bf16 output)."""
    g = lambda k: state[k]
    ob, pb = r(rg(obj, round_grads and obj_grad_bf16)), r(pred)
    x = rg(torch.cat([ob[s_idx], pb, ob[o_idx]], 1), round_grads)                       # dX
    hidden = r(torch.relu(rg(x @ r(g("net1.0.weight")).T + g("net1.0.bias"), round_grads)))      # dhid
    conf = conf_fn()
    out = r(torch.relu(rg(hidden @ r(g("net1.2.weight")).T + g("net1.2.bias"), round_grads)) * conf[:, None])
    NO = obj.shape[0]
    v = valid
    pooled = torch.zeros(NO, H, device=obj.device, dtype=out.dtype).index_add(0, s_idx[v], out[v][:, :H])
    pooled = pooled.index_add(0, o_idx[v], out[v][:, H + Dpo:])
    cnt = torch.zeros(NO, device=obj.device, dtype=conf.dtype).index_add(0, s_idx[v], conf[v]).index_add(0, o_idx[v], conf[v])
    pooled = torch.where((cnt > 0)[:, None], pooled / torch.where(cnt > 0, cnt, torch.ones_like(cnt))[:, None], pooled)
    h2 = r(torch.relu(rg(r(pooled) @ r(g("net2.0.weight")).T + g("net2.0.bias"), round_grads)))  # dh2
    new_obj = r(torch.relu(rg(h2 @ r(g("net2.2.weight")).T + g("net2.2.bias"), round_grads)))     # g4
    return new_obj, out[:, H:H + Dpo]


def model_bf16_ref(state, objs, triplets, types, padding_id, num_layers=5, H=512, D=128):
    """``Sg2LayoutModel.forward`` (sg2im/model.py:90-124, padded batch, single attribute) in the tensor-core engine's
    arithmetic: every tensor the kernels store as bf16 (embedding rows, hidden, net1 output, pooled, net2 hidden, layer
    output, box_net hidden) is rounded to bf16, all sums are fp32.  Returns (obj_vecs [B*O, D], boxes [B*O, 4])."""
    B, O, T = objs.shape[0], objs.shape[1], triplets.shape[1]
    dev = objs.device
    s, p, o = triplets[:, :, 0], triplets[:, :, 1], triplets[:, :, 2]
    base = (torch.arange(B, device=dev) * O)[:, None]
    sg, og = (s + base).reshape(-1), (o + base).reshape(-1)
    pf, tyf = p.reshape(-1), types.reshape(-1)
    obj = state["attribute_embedding.att_emb_0.weight"][objs.reshape(-1)]
    pred = state["pred_embeddings.weight"][pf]
    w_trans = state["trans_candidates_weights"]
    conf_fn = lambda: (tyf == 0).to(w_trans.dtype) + (tyf == 1).to(w_trans.dtype) * torch.sigmoid(w_trans)[pf]
    for i in range(num_layers):
        st = {k[len("gconvs.%d." % i):]: v for k, v in state.items() if k.startswith("gconvs.%d." % i)}
        obj, pred = layer_bf16_ref(st, obj, pred, sg, og, pf != padding_id, conf_fn, H, D, round_grads=True,
                                   obj_grad_bf16=True)     # the embedding rows are bf16 too: every dobj is bf16
    h = r(torch.relu(rg(r(rg(obj)) @ r(state["box_net.0.weight"]).T + state["box_net.0.bias"])))
    boxes = h @ state["box_net.2.weight"].T + state["box_net.2.bias"]
    return obj, boxes
