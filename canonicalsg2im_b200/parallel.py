"""Multi-GPU plumbing: scene graphs are independent, so a batch is cut on graph
boundaries into contiguous ranges balanced by triple count and each rank runs the whole path on its
range.  The only exchange is the sum of the weight gradients (the reference's ``nn.DataParallel``
reduce-to-GPU-0, ``sg2im/meta_models.py:17``), done here as bucketed NCCL all-reduces that start as
soon as a bucket's gradients are final, overlapping the rest of the backward pass.
"""
import torch
import torch.distributed as dist


def shard_by_cost(costs, world_size):
    """Contiguous ranges [(start, end)] * world_size over len(costs) units, balancing sum(cost):
    cut k falls where the prefix sum first reaches k/world of the total."""
    n = len(costs)
    total = float(sum(costs))
    bounds = [0]
    acc, k = 0.0, 1
    for i, c in enumerate(costs):
        acc += c
        while k < world_size and acc >= total * k / world_size:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world_size:
        bounds.append(n)
    bounds.append(n)
    bounds = [min(b, n) for b in bounds]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class BucketedGradAllReduce:
    """Flat per-bucket gradient buffers whose slices ARE the parameters' ``.grad`` (no copies);
    a bucket is all-reduced asynchronously once every parameter in it has accumulated its gradient.

    buckets: list of lists of parameters, in the order their gradients become final during
    backward (last layer first).  ``finish()`` waits for the collectives and averages."""

    def __init__(self, buckets, group=None, average=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        self.flats, self.params, self.pending, self.handles = [], [], [], []
        self._bucket_of = {}
        seen = set()
        for bi, params in enumerate(buckets):
            params = [p for p in params if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in params)
            if not params:
                self.flats.append(None)
                self.params.append([])
                self.pending.append(0)
                continue
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[id(p)] = bi
                if self.world > 1:
                    p.register_post_accumulate_grad_hook(self._hook)
            self.flats.append(flat)
            self.params.append(params)
            self.pending.append(len(params))
        self._count = list(self.pending)

    def _hook(self, p):
        bi = self._bucket_of[id(p)]
        self._count[bi] -= 1
        if self._count[bi] == 0:
            self.handles.append(dist.all_reduce(self.flats[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Wait for outstanding all-reduces; launch the ones whose hooks did not all fire (unused params)."""
        if self.world > 1:
            for bi, c in enumerate(self._count):
                if c != 0 and self.flats[bi] is not None:
                    self.handles.append(dist.all_reduce(self.flats[bi], op=dist.ReduceOp.SUM, group=self.group,
                                                        async_op=True))
            for h in self.handles:
                h.wait()
            if self.average:
                for f in self.flats:
                    if f is not None:
                        f.div_(self.world)
        self.handles = []
        self._count = list(self.pending)

    def zero(self):
        for f in self.flats:
            if f is not None:
                f.zero_()

    @property
    def nbytes(self):
        return sum(f.numel() * f.element_size() for f in self.flats if f is not None)
