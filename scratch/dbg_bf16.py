import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests import golden_inputs as gi
from tests.util import t, rel_err
from canonicalsg2im_b200.graph import GraphTripleConv

def run(prec):
    st = gi.layer_state()
    w = torch.nn.Parameter(t(st["predicates_transitive_weights"]))
    layer = GraphTripleConv(128,128,128,128,512,1,predicates_transitive_weights=w, precision=prec).cuda()
    layer.load_state_dict({k: t(v) for k, v in st.items()}, strict=True)
    obj, pred, s, o, p, ty = gi.layer_inputs()
    oo, pp = t(obj).requires_grad_(True), t(pred).requires_grad_(True)
    edges = t(np.stack([s, o], -1))
    a, b = layer(oo, pp, edges, t(p) != 0, t(ty), t(p))
    go, gp = gi.layer_out_grads(a.shape, b.shape)
    mode = sys.argv[1] if len(sys.argv) > 1 else "both"
    loss = 0
    if mode in ("both", "obj"): loss = loss + (a.float() * t(go)).sum()
    if mode in ("both", "p"): loss = loss + (b.float() * t(gp)).sum()
    loss.backward()
    out = {"new_obj": a.float(), "new_p": b.float(), "d_obj": oo.grad, "d_pred": pp.grad, "dwt": w.grad}
    for n, prm in layer.named_parameters():
        if n != "predicates_transitive_weights": out["d_" + n] = prm.grad
    return out
a = run("fp32"); b = run("bf16")
for k in a: print("%-20s %.3e" % (k, rel_err(b[k], a[k])))
