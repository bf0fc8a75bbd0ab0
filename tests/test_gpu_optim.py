"""csg_adam_multi (csrc/optim.cu) vs torch.optim.Adam on the same fp32 parameters and gradients: 1e-6 relative after
several steps (same formulas in fp32; only the association of a few multiplications differs)."""
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch(wd):
    from canonicalsg2im_b200.optim import FusedAdam
    g = torch.Generator("cuda").manual_seed(0)
    shapes = [(512, 384), (512,), (1152, 512), (1152,), (50,), (7, 3), (1,), (184, 128)] * 8     # 64 tensors: 2 launches
    base = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    # an unaligned parameter (a 4-byte offset view made contiguous in its own storage slice)
    flat = torch.randn(1001, device="cuda", generator=g)
    ours = [torch.nn.Parameter(b.clone()) for b in base] + [torch.nn.Parameter(flat[1:].clone())]
    ref = [torch.nn.Parameter(b.clone()) for b in base] + [torch.nn.Parameter(flat[1:].clone())]
    opt_a = FusedAdam(ours, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    opt_b = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    for it in range(5):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (10.0 ** (it - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        if it == 3:      # a parameter without gradient is skipped, as torch does
            ours[2].grad = None
            ref[2].grad = None
        opt_a.step()
        opt_b.step()
        opt_a.zero_grad()
        opt_b.zero_grad()
    for i, (a, b) in enumerate(zip(ours, ref)):      # includes the skipped parameter: its own step count lags by one
        assert_close(a.detach(), b.detach(), 1e-6, "param %d" % i)
    # state dicts are interchangeable with torch.optim.Adam's (scripts/train.py checkpoints)
    sd_a, sd_b = opt_a.state_dict(), opt_b.state_dict()
    assert sorted(sd_a["state"].keys()) == sorted(sd_b["state"].keys())
    for i in sd_b["state"]:
        assert float(sd_a["state"][i]["step"]) == float(sd_b["state"][i]["step"])
        assert_close(sd_a["state"][i]["exp_avg"], sd_b["state"][i]["exp_avg"], 4e-6, "exp_avg %d" % i)
    opt_c = FusedAdam(ours, lr=1.0)
    opt_c.load_state_dict(sd_b)                      # a torch.optim.Adam checkpoint loads
    for a, b in zip(ours, ref):
        gr = torch.randn(a.shape, device="cuda", generator=g)
        a.grad, b.grad = gr.clone(), gr.clone()
    opt_c.step()
    opt_b.step()
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert_close(a.detach(), b.detach(), 1e-6, "param %d after loading torch state" % i)


def test_fused_adam_rejects_cpu_params():
    from canonicalsg2im_b200.optim import FusedAdam
    with pytest.raises(RuntimeError):
        FusedAdam([torch.nn.Parameter(torch.zeros(4))])
