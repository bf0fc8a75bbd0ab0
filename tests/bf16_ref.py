"""Plain-PyTorch fp32 reference of the tensor-core engine's arithmetic: the reference layer
(sg2im/graph.py:44-113) with operands rounded to bf16 at the points where the kernels store bf16
(inputs, weights, hidden, net1 output, pooled, net2 hidden, output) and fp32 accumulation in between.
Rounding is straight-through for autograd, so its gradients have the same ReLU masks as the kernels'
and differ only by the bf16 rounding of intermediate gradients."""
import torch


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def r(x):
    return _Round.apply(x)


def layer_bf16_ref(state, obj, pred, s_idx, o_idx, valid, conf_fn, H, Dpo):
    """Flat tensors: obj [NO, Din], pred [NT, Dp], s_idx / o_idx [NT] long (global), valid [NT] bool."""
    g = lambda k: state[k]
    ob, pb = r(obj), r(pred)
    x = torch.cat([ob[s_idx], pb, ob[o_idx]], 1)
    hidden = r(torch.relu(x @ r(g("net1.0.weight")).T + g("net1.0.bias")))
    conf = conf_fn()
    out = r(torch.relu(hidden @ r(g("net1.2.weight")).T + g("net1.2.bias")) * conf[:, None])
    NO = obj.shape[0]
    v = valid
    pooled = torch.zeros(NO, H, device=obj.device).index_add(0, s_idx[v], out[v][:, :H])
    pooled = pooled.index_add(0, o_idx[v], out[v][:, H + Dpo:])
    cnt = torch.zeros(NO, device=obj.device).index_add(0, s_idx[v], conf[v]).index_add(0, o_idx[v], conf[v])
    pooled = torch.where((cnt > 0)[:, None], pooled / torch.where(cnt > 0, cnt, torch.ones_like(cnt))[:, None], pooled)
    h2 = r(torch.relu(r(pooled) @ r(g("net2.0.weight")).T + g("net2.0.bias")))
    new_obj = r(torch.relu(h2 @ r(g("net2.2.weight")).T + g("net2.2.bias")))
    return new_obj, out[:, H:H + Dpo]
