"""Multi-process (world_size 2, gloo, CPU) test of the batch-shard + gradient all-reduce logic used for multi-GPU
training (stereospike_b200/parallel.py).  The CUDA hot path itself cannot run here, so a small conv net stands in for
the model: what is checked is that sharding + GradientSynchronizer reproduce the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3, padding=1, bias=False), torch.nn.Tanh(),
                               torch.nn.Conv2d(4, 1, 3, padding=1, bias=True))


def _worker(rank, world, port, n_samples, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stereospike_b200 import parallel
    torch.manual_seed(1)
    x = torch.randn(n_samples, 2, 6, 7)
    y = torch.randn(n_samples, 1, 6, 7)
    net = _model()
    xs, ys = parallel.shard_batch(x), parallel.shard_batch(y)
    loss = (net(xs) - ys).abs().mean()            # per-rank mean loss, like the reference's per-batch mean
    loss.backward()
    sync = parallel.GradientSynchronizer(net.parameters(), bucket_bytes=64)     # tiny buckets -> several collectives
    nb = sync.sync(local_samples=xs.shape[0], global_samples=n_samples)
    q.put((rank, nb, [p.grad.clone() for p in net.parameters()], tuple(xs.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_samples', [4, 5])
def test_sharded_gradients_match_full_batch(n_samples):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: full batch, mean loss
    torch.manual_seed(1)
    x = torch.randn(n_samples, 2, 6, 7)
    y = torch.randn(n_samples, 1, 6, 7)
    net = _model()
    (net(x) - y).abs().mean().backward()
    ref = [p.grad for p in net.parameters()]
    sizes = sorted(g[3][0] for g in got)
    assert sum(sizes) == n_samples and sizes[-1] - sizes[0] <= 1
    for rank, nb, grads, _ in got:
        assert nb >= 2
        for a, b in zip(grads, ref):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_shard_bounds_cover_batch():
    from stereospike_b200 import parallel
    for n in (0, 1, 7, 8, 128):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _overlap_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stereospike_b200 import parallel

    class _Eng:            # stands in for engine.Engine: only the hook attribute is used
        grad_hook = None
    eng = _Eng()
    sync = parallel.OverlappedGradientSync(eng, small_numel=32).attach()
    sync.set_batch(local_samples=3 if rank == 0 else 1, global_samples=4)      # uneven shards
    assert eng.grad_hook is sync
    g = torch.Generator().manual_seed(10 + rank)
    grads = [torch.randn(n, generator=g) for n in (5, 4096, 7, 640, 1, 100)]   # big ones reduce in place, small ones coalesce
    sync.begin()
    for t in grads:                 # the order the engine calls the hook in: one gradient after the other
        sync.ready(t)
    sync.ready(None)                # a parameter without gradient (non-PLIF decay slot)
    sync.finish()
    q.put((rank, sync.collectives, [t.clone() for t in grads]))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_sync_reduces_in_place():
    """OverlappedGradientSync (the hook engine._run_backward calls per block): large tensors are all-reduced in place as they
    become ready, small ones in one coalesced message; result = sample-weighted mean over the ranks."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = []
    gens = [torch.Generator().manual_seed(10 + r) for r in range(world)]
    per_rank = [[torch.randn(n, generator=g) for n in (5, 4096, 7, 640, 1, 100)] for g in gens]
    for a, b in zip(*per_rank):
        want.append(a * 0.75 + b * 0.25)
    for rank, ncoll, grads in got:
        assert ncoll == 4          # 4096, 640, 100 in place + one coalesced message for 5 + 7 + 1
        for a, b in zip(grads, want):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
