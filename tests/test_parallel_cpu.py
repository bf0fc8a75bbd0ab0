"""Host logic of the multi-GPU path on CPU: the graph-boundary sharder and the bucketed gradient
all-reduce (gloo, world_size 2)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from canonicalsg2im_b200.parallel import shard_by_cost, BucketedGradAllReduce


def test_shard_by_cost_balances_and_covers():
    costs = [5, 1, 1, 1, 8, 2, 2, 4, 4, 4, 1, 7]
    for world in (1, 2, 3, 4, 8):
        ranges = shard_by_cost(costs, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == len(costs)
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        sums = [sum(costs[a:b]) for a, b in ranges]
        assert max(sums) <= sum(costs) / world + max(costs)
    assert shard_by_cost([], 2) == [(0, 0), (0, 0)]
    assert shard_by_cost([3], 4)[0] == (0, 1)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    red = BucketedGradAllReduce([list(net[2].parameters()), list(net[0].parameters())])
    x = torch.arange(24, dtype=torch.float32).view(4, 6) / 10 + rank        # each rank: its own shard
    net(x).pow(2).sum().backward()
    red.finish()
    out[rank] = [p.grad.clone() for p in net.parameters()]
    # second step: buffers are reusable after zero()
    red.zero()
    net(x).pow(2).sum().backward()
    red.finish()
    out[rank + world] = [p.grad.clone() for p in net.parameters()]
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_mean_of_shard_grads():
    world, port = 2, 29611
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref = None
    for rank in range(world):
        net.zero_grad()
        x = torch.arange(24, dtype=torch.float32).view(4, 6) / 10 + rank
        net(x).pow(2).sum().backward()
        g = [p.grad.clone() for p in net.parameters()]
        ref = g if ref is None else [a + b for a, b in zip(ref, g)]
    ref = [r / world for r in ref]
    for rank in range(world):
        for a, b in zip(out[rank], ref):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
        for a, b in zip(out[rank + world], ref):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)


class _FlatGrads(torch.autograd.Function):
    """Gradients handed to autograd as views of ONE flat buffer, as the native layer executor does."""

    @staticmethod
    def forward(ctx, x, a, b):
        ctx.save_for_backward(x)
        return (x @ a).sum() + (x.sum(0) * b).sum()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        flat = torch.empty(6 * 4 + 6)
        da, db = torch.split(flat, [24, 6])
        da.view(6, 4).copy_(x.sum(0)[:, None].expand(6, 4) * g)
        db.copy_(x.sum(0) * g)
        return None, da.view(6, 4), db


def _worker_in_place(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = torch.nn.Parameter(torch.ones(6, 4)), torch.nn.Parameter(torch.ones(6))
    red = BucketedGradAllReduce([[a, b]])
    x = torch.arange(24, dtype=torch.float32).view(4, 6) / 10 + rank
    for _ in range(2):
        _FlatGrads.apply(x, a, b).backward()
        red.finish()
        out[rank] = (a.grad.clone(), b.grad.clone(), red.in_place_buckets)
        red.zero()
    dist.destroy_process_group()


def test_bucket_reduced_in_place_on_the_producers_flat_buffer():
    world, port = 2, 29613
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_in_place, args=(world, port, out), nprocs=world, join=True)
    xs = [torch.arange(24, dtype=torch.float32).view(4, 6) / 10 + r for r in range(world)]
    db = sum(x.sum(0) for x in xs) / world
    for rank in range(world):
        ga, gb, n_in_place = out[rank]
        assert n_in_place == 1                      # no pack / unpack copies for this bucket
        assert torch.allclose(gb, db, rtol=1e-6) and torch.allclose(ga, db[:, None].expand(6, 4), rtol=1e-6)


def _worker_groups(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.arange(24, dtype=torch.float32).view(4, 6) / 10 + rank
    res = {}
    for name, groups in (("each", None), ("two", [[0, 1], [2]]), ("end", [[0, 1, 2]])):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4), torch.nn.ReLU(),
                                  torch.nn.Linear(4, 3))
        red = BucketedGradAllReduce([list(net[4].parameters()), list(net[2].parameters()), list(net[0].parameters())],
                                    launch_groups=groups)
        for _ in range(2):          # second step: the group bookkeeping is reset by finish()
            red.zero()
            net(x).pow(2).sum().backward()
            red.finish()
            res[name] = ([p.grad.clone() for p in net.parameters()], red.launches)
    out[rank] = res
    with pytest.raises(ValueError):
        BucketedGradAllReduce([[torch.nn.Parameter(torch.ones(2))], [torch.nn.Parameter(torch.ones(2))]],
                              launch_groups=[[0]])
    dist.destroy_process_group()


def test_coalesced_launch_groups_give_the_same_gradients_with_fewer_launches():
    world, port = 2, 29615
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_groups, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        res = out[rank]
        assert res["each"][1] == 3 and res["two"][1] == 2 and res["end"][1] == 1
        for name in ("two", "end"):
            for a, b in zip(res[name][0], res["each"][0]):
                assert torch.equal(a, b)
        for a, b in zip(out[0]["each"][0], res["each"][0]):
            assert torch.equal(a, b)
