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
    """Bucketed, overlapped all-reduce of the weight gradients, in place on the buffers the backward kernels wrote.

    buckets: list of lists of parameters, in the order their gradients become final during backward (last layer
    first).  A post-accumulate hook counts a bucket's gradients; once all are final the bucket is all-reduced
    asynchronously (NCCL: ``ReduceOp.AVG``, no separate scaling pass) while the remaining layers run backward:

      * if the bucket's ``.grad`` tensors tile one contiguous region of a common base buffer -- the case of a
        ``GraphTripleConv`` layer, whose executor writes all eight weight / bias gradients into one flat buffer
        (``graph_tc._TripleConvEngine.backward``) -- that region itself is reduced: no flat copy, no accumulate
        pass, nothing to zero between steps;
      * otherwise the gradients are packed into a temporary flat buffer, reduced, and unpacked in ``finish()``.

    ``finish()`` waits for the collectives; ``zero()`` drops the gradients (``set_to_none``)."""

    def __init__(self, buckets, group=None, average=True, launch_groups=None):
        """``launch_groups``: lists of bucket indices whose all-reduces are issued TOGETHER, as one coalesced NCCL
        launch (``ncclGroupStart`` ... ``ncclGroupEnd``: one kernel for all regions), once the last of them is final.
        ``None``: every bucket on its own.  Why one might group: the kernels of the backward pass are persistent and
        fill every SM, so an all-reduce kernel runs between two of them rather than beside them and each launch puts
        its latency floor on the critical path.  Measured: worth 2 % of a cfg2 step at N = 2 and
        nothing at N = 8, so ``SgToLayoutStep`` keeps one launch per bucket unless ``CSG_GRAD_GROUPS`` says otherwise."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        backend = dist.get_backend(group) if dist.is_initialized() else ""
        self.native_avg = average and backend == "nccl"
        self.params, self.pending = [], []
        self._bucket_of = {}
        self._inflight = []          # (handle, region, unpack list or None)
        # Stream a bucket's all-reduce is launched from.  Autograd runs a node's backward on the stream of its forward,
        # so a gradient of a bucket may become final on another stream than its neighbours (the canvas branch of
        # pipeline.SgToLayoutStep): the hook then leaves an event there and the launch makes `home` wait for it.
        self.home = None
        self._events = [[] for _ in buckets]
        self.in_place_buckets = 0    # statistics of the last step (tests / bench)
        self.launches = 0            # collective launches of the last step
        seen = set()
        for bi, params in enumerate(buckets):
            params = [p for p in params if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in params)
            for p in params:
                self._bucket_of[id(p)] = bi
                if self.world > 1:
                    p.register_post_accumulate_grad_hook(self._hook)
            self.params.append(params)
            self.pending.append(len(params))
        self._count = list(self.pending)
        if launch_groups is None:
            launch_groups = [[bi] for bi in range(len(buckets))]
        covered = sorted(bi for g in launch_groups for bi in g)
        if covered != list(range(len(buckets))):
            raise ValueError("launch_groups must cover every bucket exactly once")
        self.launch_groups = [list(g) for g in launch_groups]
        self._group_of = {bi: gi for gi, g in enumerate(self.launch_groups) for bi in g}
        self._group_left = [sum(1 for bi in g if self.params[bi]) for g in self.launch_groups]
        self._group_pending = list(self._group_left)
        self._launched = [False] * len(self.launch_groups)

    @staticmethod
    def _contiguous_region(grads):
        """The 1-D tensor covering exactly the given gradients if they tile one region of a common storage
        (autograd hands over the producer's views, so ``.grad`` shares the storage of the flat buffer the backward
        kernels wrote)."""
        g0 = grads[0]
        st = g0.untyped_storage()
        esz = g0.element_size()
        spans = []
        for g in grads:
            if (g.untyped_storage().data_ptr() != st.data_ptr() or not g.is_contiguous() or g.dtype != g0.dtype
                    or g.device != g0.device):
                return None
            spans.append((g.data_ptr(), g.data_ptr() + g.numel() * esz))
        spans.sort()
        for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
            if a1 != b0:
                return None
        off = (spans[0][0] - st.data_ptr()) // esz
        n = (spans[-1][1] - spans[0][0]) // esz
        return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, off, (n,), (1,))

    def _launch(self, gi):
        """Issue the all-reduces of launch group ``gi`` (all of its buckets that have gradients) as one launch."""
        self._launched[gi] = True
        items = []
        for bi in self.launch_groups[gi]:
            grads = [p.grad for p in self.params[bi] if p.grad is not None]
            if grads:
                items.append((bi, grads))
        if not items:
            return
        if self.home is not None and items[0][1][0].is_cuda:
            # everything below (packing included) is issued on `home`, behind the gradients that became final elsewhere
            for bi, _ in items:
                for ev in self._events[bi]:
                    self.home.wait_event(ev)
            with torch.cuda.stream(self.home):
                self._launch_on_current(items)
        else:
            self._launch_on_current(items)
        for bi, _ in items:
            self._events[bi] = []

    def _launch_on_current(self, items):
        op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
        regions, unpacks = [], []
        for _, grads in items:
            region = self._contiguous_region(grads)
            unpack = None
            if region is None:
                region = torch.cat([g.reshape(-1) for g in grads])
                unpack = grads
            else:
                self.in_place_buckets += 1
            regions.append(region)
            unpacks.append(unpack)
        if len(regions) == 1:
            h = dist.all_reduce(regions[0], op=op, group=self.group, async_op=True)
        else:
            pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
            opts = dist.AllreduceCoalescedOptions()
            opts.reduceOp = op
            opts.asyncOp = True
            h = pg.allreduce_coalesced(regions, opts)
        self.launches += 1
        for region, unpack in zip(regions, unpacks):
            self._inflight.append((h, region, unpack))
            h = None          # one handle per launch: waited once

    def _hook(self, p):
        bi = self._bucket_of[id(p)]
        if self.home is not None and p.is_cuda:
            cur = torch.cuda.current_stream()
            if cur != self.home:
                ev = torch.cuda.Event()
                ev.record(cur)
                self._events[bi].append(ev)
        self._count[bi] -= 1
        if self._count[bi] == 0:
            gi = self._group_of[bi]
            self._group_pending[gi] -= 1
            if self._group_pending[gi] == 0:
                self._launch(gi)

    def finish(self):
        """Wait for outstanding all-reduces; launch the ones whose hooks did not all fire (unused params)."""
        if self.world > 1:
            for gi, done in enumerate(self._launched):
                if not done and self._group_left[gi]:
                    self._launch(gi)
            for h, region, unpack in self._inflight:
                if h is not None:
                    h.wait()
                if self.average and not self.native_avg:
                    region.div_(self.world)
                if unpack is not None:
                    off = 0
                    for g in unpack:
                        g.copy_(region[off:off + g.numel()].view_as(g))
                        off += g.numel()
        self._inflight = []
        self._count = list(self.pending)
        self._group_pending = list(self._group_left)
        self._launched = [False] * len(self.launch_groups)

    def zero(self):
        self.in_place_buckets = 0
        self.launches = 0
        for params in self.params:
            for p in params:
                p.grad = None

    @property
    def nbytes(self):
        return sum(p.numel() * p.element_size() for params in self.params for p in params)
