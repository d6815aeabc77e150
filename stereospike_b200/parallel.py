"""Batch-sharded multi-GPU execution: one process per GPU (``torch.distributed``, NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The hot path shards on the batch only -- no BatchNorm, no cross-sample term (SURVEY.md
section 8(e)) -- so inference is N independent replicas with no data-path collective, and training needs exactly one
collective per step: a sum all-reduce of the 18.15 M-parameter gradient (72.6 MB fp32), issued in a few large buckets.
The reference has no distributed code; this mirrors its single-process step (train.py:189-242) per rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world_size):
    """[lo, hi) of rank's contiguous shard of n samples (shards differ by at most one sample)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world_size=None, dim=0):
    """This rank's contiguous slice of ``x`` along the batch dimension (a view; no communication)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(x.shape[dim], rank, world_size)
    return x.narrow(dim, lo, hi - lo)


class GradientSynchronizer:
    """Flat, pre-allocated gradient buckets: after ``backward()`` call ``sync()`` -- gradients are copied into the
    buckets, all-reduced (sum) asynchronously bucket by bucket in reverse parameter order, averaged over the global
    batch and copied back.  ``weights`` (per-rank sample counts) makes the average exact for uneven shards."""

    def __init__(self, params, bucket_bytes=32 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets = []          # (flat tensor, [(param, offset, numel)])
        cur, cur_n = [], 0
        for p in reversed(self.params):
            n = p.numel()
            if cur and (cur_n + n) * 4 > bucket_bytes:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append((p, cur_n, n))
            cur_n += n
        if cur:
            self._close(cur, cur_n)

    def _close(self, items, n):
        dev = items[0][0].device
        self.buckets.append((torch.zeros(n, dtype=torch.float32, device=dev), items))

    def sync(self, local_samples=1, global_samples=None):
        """Turns per-rank gradients of a per-rank MEAN loss into the gradient of the global-batch mean loss."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return 0
        world = dist.get_world_size(self.group)
        if global_samples is None:
            global_samples = local_samples * world
        scale = float(local_samples) / float(global_samples)
        works = []
        for flat, items in self.buckets:
            for p, off, n in items:
                g = p.grad
                if g is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(g.reshape(-1))
            flat.mul_(scale)
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for (flat, items), w in zip(self.buckets, works):
            w.wait()
            for p, off, n in items:
                if p.grad is None:
                    p.grad = flat[off:off + n].reshape(p.shape).clone()
                else:
                    p.grad.copy_(flat[off:off + n].reshape(p.shape))
        return len(self.buckets)


class OverlappedGradientSync:
    """Gradient all-reduce issued from INSIDE the backward pass, block by block ("bucketed in
    reverse-layer order and overlapped with the remaining backward"; the step it mirrors per rank: train.py:189-242).

    The engine computes the blocks' weight gradients in reverse layer order (engine._run_backward).  As soon as the kernel that
    produces one is enqueued, ``ready(g)`` launches ``dist.all_reduce(g, async_op=True)`` IN PLACE on that tensor: NCCL's stream
    waits (event) for the producing kernel and then runs concurrently with the dgrad / wgrad kernels of the earlier layers on
    the compute stream.  No flat copy-in / copy-out of the 72.6 MB gradient; only the tiny tensors (head weights / biases, PLIF
    decays: < ``small_numel`` elements each) are coalesced into one flat message at the end.  ``finish()`` makes the compute
    stream wait for the outstanding collectives before autograd hands the gradients to the optimizer -- by then only the last
    (smallest: the first layer's 3 200 weights) all-reduce can still be in flight, so almost nothing is exposed.

        sync = OverlappedGradientSync(net).attach()
        sync.set_batch(local_samples=B, global_samples=B * world)
        loss.backward()          # all-reduces happen in here
        optimizer.step()

    Gradients of a per-rank MEAN loss are scaled by local/global samples and summed -> gradient of the global-batch mean."""

    def __init__(self, net, group=None, small_numel=1 << 16):
        self.engine = net.engine if hasattr(net, 'engine') else net
        self.group = group
        self.small_numel = int(small_numel)
        self.scale = None
        self._works, self._small = [], []
        self.collectives = 0          # issued during the last backward (tests / bench)

    def attach(self):
        self.engine.grad_hook = self
        return self

    def detach(self):
        if self.engine.grad_hook is self:
            self.engine.grad_hook = None
        return self

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def set_batch(self, local_samples=1, global_samples=None):
        world = dist.get_world_size(self.group) if self._active() else 1
        if global_samples is None:
            global_samples = local_samples * world
        self.scale = float(local_samples) / float(global_samples)
        return self

    def begin(self):
        self._works, self._small, self.collectives = [], [], 0

    def ready(self, g):
        """``g``: a finished (enqueued on the current stream) contiguous gradient tensor; reduced in place."""
        if g is None or not self._active():
            return
        if g.numel() < self.small_numel:
            self._small.append(g)
            return
        scale = self.scale if self.scale is not None else 1.0 / dist.get_world_size(self.group)
        g.mul_(scale)
        self._works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.collectives += 1

    def finish(self):
        if not self._active():
            return
        if self._small:
            scale = self.scale if self.scale is not None else 1.0 / dist.get_world_size(self.group)
            flat = torch.cat([g.reshape(-1) for g in self._small]).mul_(scale)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()
            self.collectives += 1
            off = 0
            for g in self._small:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        for w in self._works:
            w.wait()              # the compute stream waits for NCCL's stream; the host does not block
        self._works, self._small = [], []


def allreduce_gradients(module_or_params, local_samples=1, global_samples=None, bucket_bytes=32 << 20, group=None):
    """One-shot helper around :class:`GradientSynchronizer` (allocates the buckets on every call)."""
    params = module_or_params.parameters() if hasattr(module_or_params, 'parameters') else module_or_params
    return GradientSynchronizer(list(params), bucket_bytes, group).sync(local_samples, global_samples)
