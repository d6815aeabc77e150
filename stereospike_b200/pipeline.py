"""Host -> device streaming of event-frame batches for the fused forward path.

The reference moves every batch with a blocking ``.to(device)`` before it calls the network (train.py:193-198,
test.py:113-118).  Here the copy of batch i+1 runs on a second CUDA stream while batch i is being computed, and the
finest depth map is copied back asynchronously, so that the PCIe transfer is hidden behind the kernels:

    pipe = HostPipeline(net, batch_shape=(8, 5, 4, 260, 346))
    for depth in pipe.run(batches):          # batches: iterable of pinned fp32 host tensors [B,T,C,H,W]
        ...                                   # depth: pinned host tensor [B,1,H,W], valid after pipe.sync() (or pipe.done[k])
"""
import torch

from . import functional


def pack_events_host(x_btchw):
    """Host-side data preparation (NOT the hot path): the reference's fp32 count frames [B,T,C,H,W] (C <= 4) as the packed
    u8 [T,B,H,W,4] layout the first block reads -- 4x fewer bytes over PCIe than the fp32 frames.  A data loader that
    histograms events straight into u8 counts (events.cumulate_spikes_into_frames does it on the device) never needs this."""
    assert x_btchw.dim() == 5 and x_btchw.shape[2] <= 4 and not x_btchw.is_cuda
    B, T, C, H, W = x_btchw.shape
    out = torch.zeros((T, B, H, W, 4), dtype=torch.uint8)
    out[..., :C] = x_btchw.permute(1, 0, 3, 4, 2).round().clamp(0, 255).to(torch.uint8)
    return out


class HostPipeline:
    """``batch_shape``/``dtype``: fp32 ``[B,T,C,H,W]`` frames (the reference's tensors) or packed u8 ``[T,B,H,W,4]`` count
    frames (``pack_events_host`` / a u8 data loader: 14.4 MB instead of 57.6 MB per B=8, T=5 batch).  ``stateless``: every step
    starts from reset neurons (what test.py does per sample), so the final membrane potentials are never read and the blocks
    skip writing them (``keep_state=False``: -377 MB of HBM writes per B=8, T=5 step)."""

    def __init__(self, net, batch_shape, device=None, depth=2, dtype=torch.float32, stateless=True):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.device)
        self.bufs = [torch.empty(batch_shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        if dtype == torch.uint8:
            _, B, H, W, _ = batch_shape
        else:
            B, _, _, H, W = batch_shape
        self.depth_host = [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(depth)]
        # the depth map goes back on a stream of its own: on the compute stream its 2.9 MB (B = 8) would hold back the next
        # batch's kernels for the 50 us PCIe takes
        self.out_stream = torch.cuda.Stream(self.device)
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self._keep = [None] * depth             # the device depth map of the slot's batch, alive until its copy has finished
        self.stateless = bool(stateless)
        self._i = 0

    def step(self, x_host):
        """Enqueue one batch (pinned host tensor); returns the pinned host depth map it will land in."""
        k = self._i % len(self.bufs)
        self._i += 1
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[k])          # the kernels that read this buffer have finished
            self.bufs[k].copy_(x_host, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        main.wait_event(self.ready[k])
        functional.reset_net(self.net)
        eng = self.net.engine
        keep = eng.keep_state
        if self.stateless:
            eng.keep_state = False
        try:
            with torch.no_grad():
                out = self.net.forward_seq(self.bufs[k])
        finally:
            eng.keep_state = keep
        self.free[k].record(main)
        depths = out if not isinstance(out, tuple) else out[0]
        if self._keep[k] is not None:
            main.wait_event(self.done[k])       # the slot's previous copy-out (long finished) before its tensor may be reused
        self._keep[k] = depths[0]
        self.computed[k].record(main)
        self.out_stream.wait_event(self.computed[k])
        with torch.cuda.stream(self.out_stream):
            self.depth_host[k].copy_(depths[0], non_blocking=True)
            self.done[k].record(self.out_stream)
        return self.depth_host[k]

    def join(self):
        """The current stream waits (on the device, no host synchronisation) for every depth map enqueued so far."""
        torch.cuda.current_stream(self.device).wait_stream(self.out_stream)

    def run(self, batches):
        for x in batches:
            yield self.step(x)

    def sync(self):
        torch.cuda.synchronize(self.device)


class GraphedInference:
    """Stateless sequence inference (``reset_net`` + ``forward_seq``, what test.py does per sample) captured once in a
    CUDA graph and replayed: one graph launch instead of ~20 kernel launches and their Python dispatch, which is what
    bounds small batches (B = 1, T = 1: the live event-camera case).

        run = GraphedInference(net, batch_shape=(1, 5, 4, 260, 346))
        depths = run(x)          # list [depth1..depth4] of static fp32 [B,1,H,W] tensors, overwritten by the next call

    The weights are baked into the captured launches' packed images: re-create the object after changing them.
    """

    def __init__(self, net, batch_shape, dtype=torch.float32, device=None, warmup=2):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.x = torch.zeros(batch_shape, dtype=dtype, device=self.device)
        self.stream = torch.cuda.Stream(self.device)
        eng = net.engine
        keep = eng.keep_state
        eng.keep_state = False            # nothing reads the final potentials of a stateless run: skip writing them
        try:
            cur = torch.cuda.current_stream(self.device)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):   # lazy one-time work (weight images, geometry tables, function attributes) stays outside
                    self._eager()
            cur.wait_stream(self.stream)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = self._eager()
        finally:
            eng.keep_state = keep
        functional.reset_net(net)         # the captured run left graph-owned tensors in the neuron modules

    def _eager(self):
        functional.reset_net(self.net)
        with torch.no_grad():
            out = self.net.forward_seq(self.x)
        return list(out[0]) if isinstance(out, tuple) else list(out)

    def __call__(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out
