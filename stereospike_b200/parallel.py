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


def allreduce_gradients(module_or_params, local_samples=1, global_samples=None, bucket_bytes=32 << 20, group=None):
    """One-shot helper around :class:`GradientSynchronizer` (allocates the buckets on every call)."""
    params = module_or_params.parameters() if hasattr(module_or_params, 'parameters') else module_or_params
    return GradientSynchronizer(list(params), bucket_bytes, group).sync(local_samples, global_samples)
