"""Execution engine of the spiking U-Net: walks the model's fused sites layer by layer (block l for all T
timesteps, then block l+1 -- mathematically identical to the reference's timestep loop because the network is
feed-forward in depth) and calls one CUDA kernel per block through the C ABI.

Forward : ss_pack_events + ss_conv_i8_fwd (tcgen05 int8 tensor-core kernel) per spiking block -- or ss_conv_neuron_fwd
          (fp32 CUDA cores) when impl='simt' -- then ss_heads_fwd for the four heads + I-neurons.
Backward: ss_heads_bwd, then per block in reverse order ss_neuron_bwd_ex (surrogate BPTT scan, bf16 gradient out),
          ss_conv_wgrad_bf16 and ss_corr_bf16 (tcgen05 bf16 tensor-core kernels; bwd_impl='simt' selects the fp32
          CUDA-core ss_conv_wgrad / ss_conv_dgrad).  Replaces PyTorch autograd through the reference modules
          (SURVEY.md section 3(C)).
"""
import ctypes
import os

import torch

from . import _lib, ops
from ._lib import SS_IMPL_AUTO, SS_IMPL_SIMT, SS_IMPL_UMMA, SS_IN_U8_TBHWC, SS_IN_F32_BTCHW
from .ops import BlockGeom, _ptr, _stream, conv_out_size

IMPLS = {'auto': SS_IMPL_AUTO, 'simt': SS_IMPL_SIMT, 'umma': SS_IMPL_UMMA}


class Site:
    """One fused spiking block: conv (plain or NN-upsampled) -> gain -> neuron [-> + residual]."""

    def __init__(self, name, out, src, conv, gain_mod, node, resid=None, up_size=None):
        self.name, self.out, self.src, self.resid = name, out, src, resid
        self.conv, self.gain_mod, self.node, self.up_size = conv, gain_mod, node, up_size
        self._pack = None
        self._dgrad = None

    def dgrad_plan(self, geom):
        """Correlation weights / maps of this block's data gradient (ops.DgradPlan), cached on the weight's version."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version, geom.Hin, geom.Win, geom.Hout, geom.Wout, str(w.device))
        if self._dgrad is None or self._dgrad[0] != key:
            self._dgrad = (key, ops.DgradPlan(w, geom, w.device))
        return self._dgrad[1]

    def unfolded_fold_taps(self, planes, wexp):
        """25-tap image of the taps quantised with the fold's exponents (small calls of a folded block), cached on the weight."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version, planes, str(w.device))
        if getattr(self, '_unfolded', None) is None or self._unfolded[0] != key:
            self._unfolded = (key, ops.pack_weights_with_exponents(w, planes, wexp))
        return self._unfolded[1]

    def geom(self, Hin, Win):
        c = self.conv
        ks = c.kernel_size[0]
        if self.up_size is not None:
            return BlockGeom('upconv', c.in_channels, c.out_channels, ks, Hin, Win, self.up_size[0], self.up_size[1])
        st, pd = c.stride[0], c.padding[0]
        return BlockGeom('conv', c.in_channels, c.out_channels, ks, Hin, Win, conv_out_size(Hin, ks, st, pd),
                         conv_out_size(Win, ks, st, pd), st, pd)

    def packed(self, planes, need_i8, need_kn=True, fold=False):
        """(w_kn fp32 [K][Cout] or None, (w_i8, wscale) or (dense, rows, cols, wscale, wexp) or None), cached on the weight's
        version."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version, planes, need_i8, need_kn, fold, str(w.device))
        if self._pack is None or self._pack[0] != key:
            w_kn = ops.weight_to_kn(w) if need_kn else None
            w_i8 = None
            if need_i8 and fold:
                w_i8 = ops.pack_weights_folded(w, planes, with_exp=True)          # (dense, rows, cols, wscale, wexp)
            elif need_i8:
                cin = w.shape[1]
                q, sc, _ = ops.pack_weights_i8(w, planes, cin_pad=ops.first_layer_channels(cin) if cin % 32 else cin)
                w_i8 = (q, sc)
            self._pack = (key, w_kn, w_i8)
        return self._pack[1], self._pack[2]


class Head:
    def __init__(self, name, src, upconv, gain_mod):
        self.name, self.src, self.upconv, self.gain_mod = name, src, upconv, gain_mod

    @property
    def conv(self):
        return self.upconv.up[1]

    def geom(self, Hin, Win):
        c = self.conv
        return BlockGeom('upconv', c.in_channels, 1, c.kernel_size[0], Hin, Win, self.upconv.up_size[0],
                         self.upconv.up_size[1])


def _node_v_in(node, shape_bhwc, device):
    """Membrane potential carried from a previous call as fp32 [B,H,W,C], or None when at v_reset."""
    v = node.v
    if not isinstance(v, torch.Tensor):
        if float(v) == node.v_reset:
            return None
        return torch.full(shape_bhwc, float(v), dtype=torch.float32, device=device)
    if v.dim() == 4:
        v = v.permute(0, 2, 3, 1)
    v = v.detach().to(device=device, dtype=torch.float32).contiguous()
    assert tuple(v.shape) == tuple(shape_bhwc), (tuple(v.shape), tuple(shape_bhwc))
    return v


class _NetFunction(torch.autograd.Function):
    """Whole-network autograd node: inputs are the parameter tensors, outputs are the depth stack [4,B,H,W] and, when
    asked for (side['spike_outputs']), the last-timestep spike maps of the named layers as fp32 NCHW tensors -- so that a
    loss on the spikes (the reference's SpikePenalization_Loss, loss.py:96-107) back-propagates through the surrogates."""

    @staticmethod
    def forward(ctx, eng, x_seq, side, n_site_params, *params):
        need_grad = side['need_grad']
        res = eng._run_forward(x_seq, params, n_site_params, side, want_h=need_grad)
        names = side.get('spike_outputs') or ()
        spks = tuple(side['acts'][k][-1].permute(0, 3, 1, 2).float() for k in names) if need_grad else ()
        if names and need_grad and side.get('stats') is not None:
            # SpikePenalization_Loss (loss.py:96-107) of the returned maps from the counters the block epilogues accumulated:
            # sum_k  sum(s_k^2) / (2 numel_k)  over the last timestep -- no pass over the spike maps
            pen = side['stats'].new_zeros((), dtype=torch.float64)
            for k in names:
                pen = pen + side['stats'][eng.site_index_of_output(k), 5].double() / (2.0 * side['acts'][k][-1].numel())
            spks = spks + (pen.float(),)
        depths = res['depths']
        if need_grad:
            ctx.eng, ctx.n_site_params = eng, n_site_params
            ctx.params = params
            # the outputs must not be reachable from ctx (output -> grad_fn -> ctx -> output is a cycle only the cyclic GC
            # breaks: every layer's h_seq would stay alive after a grad-enabled forward that is never back-propagated)
            ctx.saved = {k: v for k, v in res.items() if k != 'depths'}
            ctx.depth_shape, ctx.depth_device = tuple(depths.shape), depths.device
            ctx.spk_names = names
        side['spikes_fp32'] = spks
        return (depths,) + spks

    @staticmethod
    def backward(ctx, g_depths, *g_spks):
        saved = ctx.saved
        if saved is None:
            raise RuntimeError('stereospike_b200: backward through the same forward twice (the saved potentials were released)')
        if g_depths is None:
            g_depths = torch.zeros(ctx.depth_shape, dtype=torch.float32, device=ctx.depth_device)
        inject = {k: gs for k, gs in zip(ctx.spk_names, g_spks) if gs is not None}
        g_pen = g_spks[len(ctx.spk_names)] if len(g_spks) > len(ctx.spk_names) else None
        if g_pen is not None:
            # d/ds of sum(s^2) / (2 n) = s / n on the last timestep of every penalised layer (device-side, no sync)
            acts = saved['acts']
            for k in ctx.spk_names:
                s_last = acts[k][-1].permute(0, 3, 1, 2).float()
                term = s_last * (g_pen.float() / s_last.numel())
                inject[k] = term if k not in inject else inject[k] + term
        grads = ctx.eng._run_backward(saved, ctx.params, ctx.n_site_params, g_depths.contiguous().float(), inject)
        ctx.saved = None
        return (None, None, None, None) + tuple(grads)


class Engine:
    def __init__(self, sites, heads, ineuron):
        self.sites, self.heads, self.ineuron = sites, heads, ineuron
        self.impl = 'auto'
        self.weight_planes = 3
        self.keep_state = True
        self.timing = None          # bench.py: list of (site name, start event, end event) when not None
        self.check_input = True     # fp32 frames must hold integer event counts 0..255 (the tensor-core path packs them to u8): a
        #                             status word is written by ss_pack_events and checked WITHOUT a sync -- blocking on the very first
        #                             call, then through a pinned host copy inspected at the start of a later call (ValueError)
        self._status = None         # (device int32[1], pinned host int32[1], event) per device
        self._status_checked_once = False
        self.collect_stats = False  # every block adds {spikes, nonzero outputs, sum out^2} (all steps / last step) to side['stats'] [n_sites, 6]
        #                             from its epilogue registers (dp4a + one atomic per warp); always on for a grad-enabled
        #                             forward_seq(spikes_fp32=True) (fused spike penalty) and for calculate_firing_rates
        self.grad_hook = None       # parallel.OverlappedGradientSync: called with each weight-gradient tensor as soon as it is
        #                             enqueued (reverse layer order), so that its all-reduce overlaps the rest of the backward
        self.fold_upsample = True   # NNConvUpsampling blocks folded: four 3x3 convs on the source for the regular outputs + two small
        #                             passes for the irregular rows / columns (9 / 15 taps instead of 25, bit-identical integers, 3-4 bits
        #                             less weight precision).  True / False, or a collection of site names ('deconv4', ...)
        self.fold_min_frames = 16   # ... for calls of at least this many event frames (B * T); smaller calls are launch-latency-bound
        self.batch_as_steps = os.environ.get('SS_BATCH_AS_STEPS', '1') != '0'   # stateless single-step calls on a batch: k independent
        #                             "steps" of B / k samples per launch (one weight stream per tile serves k patches), bit-identical
        self.bas_min_batch = int(os.environ.get('SS_BAS_MIN_BATCH', '4'))
        self.bas_min_cin = int(os.environ.get('SS_BAS_MIN_CIN', '128'))
        self.flop_scale = {}        # site -> executed taps / 25 of the folded blocks of the last forward (bench.py credits these FLOPs)
        self.bwd_impl = 'umma'      # gradients of the convs: 'umma' = bf16 tensor cores (fp32 accumulation), 'simt' = fp32 CUDA cores
        self.wgrad_stream = os.environ.get('SS_WGRAD_STREAM', '1') != '0'    # tensor-core weight gradients on a second stream, overlapping the rest of the backward
        self._side = None
        self.heads_time_sum = True  # fold the time loop of the (linear, non-firing) readout: 2 head passes instead of T;
        #                             False = per-timestep accumulation in the reference's order (bit-identical to T single steps)

    def site_index_of_output(self, out_name):
        for i, s in enumerate(self.sites):
            if s.out == out_name:
                return i
        raise KeyError(out_name)

    def _fold_site(self, name):
        f = self.fold_upsample
        return (name in f) if isinstance(f, (set, frozenset, list, tuple)) else bool(f)

    # ------------------------------------------------------------------ parameter flattening
    def _flat_params(self):
        ps = []
        for s in self.sites:
            ps.append(s.conv.weight)
            ps.append(s.node.decay_tensor())       # None unless PLIF (sigmoid(w); torch differentiates it)
        n_site = len(ps)
        for h in self.heads:
            ps.append(h.conv.weight)
            ps.append(h.conv.bias)
        return ps, n_site

    # ------------------------------------------------------------------ public entry
    def run(self, x_seq, return_layers=False, spike_outputs=None):
        """x_seq fp32 [B,T,C,H,W] on CUDA.  Returns (depth stack [4,B,H,W] in execution order, side dict).
        ``spike_outputs``: layer names whose last-timestep spike maps should be differentiable outputs (side['spikes_fp32'])."""
        ops._require_cuda(x_seq, 'x')
        if x_seq.dim() != 5:
            raise ValueError('expected x of shape [B, T, C, H, W]')
        if x_seq.dtype == torch.uint8:
            # packed event frames u8 [T, B, H, W, 4] (stereospike_b200.events / ss_pack_events): inference only
            if x_seq.shape[-1] != ops.first_layer_channels(self.sites[0].conv.in_channels) or IMPLS[self.impl] == SS_IMPL_SIMT:
                raise ValueError('packed input must be u8 [T, B, H, W, 4] (32 per 32 channels in the channel-concatenated mode) '
                                 'and needs the tensor-core path')
            x_seq = x_seq.contiguous()
        else:
            x_seq = x_seq.contiguous().float()
        params, n_site = self._flat_params()
        need_grad = torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in params)
        self._poll_input_status()
        if need_grad:
            self._warn_truncated_bptt()
        if need_grad and x_seq.dtype == torch.uint8:
            raise NotImplementedError('packed u8 input is forward-only (the first layer\'s weight gradient reads the fp32 frames)')
        side = {'need_grad': need_grad, 'return_layers': return_layers, 'spike_outputs': tuple(spike_outputs or ())}
        outs = _NetFunction.apply(self, x_seq, side, n_site, *params)
        side['spikes_fp32'] = outs[1:]
        return outs[0], side

    # ------------------------------------------------------------------ input validation without a sync
    def _status_word(self, dev):
        if self._status is None or self._status[0].device != dev:
            self._status = (torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32).pin_memory(),
                            torch.cuda.Event())
            self._status_pending = False
        return self._status[0]

    def _poll_input_status(self, block=False):
        """Raises ValueError if an earlier call was fed fp32 frames that are not integer event counts in 0..255 (they would have
        been rounded / clamped by the u8 packing).  Non-blocking unless ``block``: reads a pinned copy once its event has fired."""
        if self._status is None or not self._status_pending or torch.cuda.is_current_stream_capturing():
            return
        _, host, ev = self._status
        if block:
            ev.synchronize()
        elif not ev.query():
            return
        self._status_pending = False
        if int(host[0]) != 0:
            self._status[0].zero_()
            raise ValueError('stereospike_b200: an input frame passed to the tensor-core path was not integer event counts in 0..255 '
                             '(normalised or > 255 events per pixel); its values were rounded / clamped. Use '
                             "set_kernel_options(impl='simt') for arbitrary fp32 input.")

    def _record_input_status(self):
        dev_w, host, ev = self._status
        host.copy_(dev_w, non_blocking=True)
        ev.record()
        self._status_pending = True
        if not self._status_checked_once:
            self._status_checked_once = True
            self._poll_input_status(block=True)        # the first call pays one sync so that a wrong data pipeline fails at once

    def _warn_truncated_bptt(self):
        """The whole network is one autograd node per call and the carried membrane potentials are detached, so several
        grad-enabled forward() calls before one backward() give TRUNCATED BPTT (the reference back-propagates through every call
        until net.detach()).  Multi-step BPTT goes through forward_seq.  Warn once instead of silently differing."""
        import warnings
        for s in self.sites:
            if isinstance(s.node.v, torch.Tensor) and getattr(s.node, '_v_from_grad_call', False):
                warnings.warn('stereospike_b200: neuron state carried in from a previous grad-enabled call is detached -- gradients '
                              'do not flow across separate forward() calls (truncated BPTT). Use forward_seq(x_seq) for multi-step '
                              'BPTT, or call net.detach() / functional.reset_net(net) between steps to silence this.', stacklevel=4)
                for t in self.sites:
                    t.node._v_from_grad_call = False
                return

    # ------------------------------------------------------------------ forward
    def _run_forward(self, x_seq, params, n_site, side, want_h):
        packed_in = x_seq.dtype == torch.uint8
        B, T = (int(x_seq.shape[1]), int(x_seq.shape[0])) if packed_in else (int(x_seq.shape[0]), int(x_seq.shape[1]))
        dev = x_seq.device
        impl = IMPLS[self.impl]
        acts = {'x': x_seq}
        acts.update(side.get('seed_acts') or {})       # stand-alone blocks: the block input under its own name (u8 [T,B,H,W,C])
        saved = {'B': B, 'T': T, 'sites': [], 'acts': acts}
        head_srcs = {h.src for h in self.heads}
        tsums = {}
        stats = None
        if (self.collect_stats or (want_h and side.get('spike_outputs'))) and impl != SS_IMPL_SIMT:
            stats = torch.zeros((len(self.sites), 6), dtype=torch.int64, device=dev)
        side['stats'] = stats
        # A single-step call on a batch (the reference's calling convention: one forward(x) per frame) is issued as k independent
        # "steps" of B / k samples: every tile then streams its weights once for k patches instead of once per patch -- the deep
        # blocks of a T = 1 call are bound by re-streaming their weights from L2.  Stateless inference only; bit-identical.
        k_steps = 1
        if self.batch_as_steps and T == 1 and B > 1 and not want_h and not self.keep_state and impl != SS_IMPL_SIMT and \
                self.weight_planes == 3 and stats is None and not side.get('seed_acts') and \
                all(not isinstance(s.node.v, torch.Tensor) and float(s.node.v) == 0.0 and s.node.v_reset == 0.0 for s in self.sites):
            # at least 3 steps of at least 4 samples (measured, T = 1: B = 16 as 4 x 4 +16 %, B = 32 as 4 x 8 +37 %; B = 8 as 2 x 4
            # -2 %, B = 5 as 5 x 1 -13 %: too few items per launch for the 148 SMs)
            k_steps = next((c for c in (5, 4, 3) if B % c == 0 and B // c >= self.bas_min_batch), 1)
        k_all = self.last_batch_steps = k_steps       # (read by the tests)
        for i, s in enumerate(self.sites):
            xin = acts[s.src]
            first = s.src == 'x'
            Hin, Win = (int(xin.shape[3]), int(xin.shape[4])) if (first and not packed_in) else (int(xin.shape[2]), int(xin.shape[3]))
            g = s.geom(Hin, Win)
            if first and not packed_in and int(x_seq.shape[2]) != g.Cin:
                raise ValueError(f'input has {int(x_seq.shape[2])} channels, the model expects {g.Cin}')
            use_i8 = impl != SS_IMPL_SIMT
            # (also while training: ss_pack_weights_folded re-derives the folded sets from the updated weight in one launch)
            fold_w = use_i8 and self._fold_site(s.name) and self.weight_planes <= 3 and g.kind == 'upconv' and g.ks == 5 and \
                g.Cin % 32 == 0 and ops.fold_plan(g.Hin, g.Win, g.Hout, g.Wout, B, str(dev)).ok
            # (not for a handful of frames: the two extra row-list launches per block cost more than the taps they save -- single-frame
            #  graph replay 0.28 ms folded vs 0.24 ms unfolded.  Such calls run the 25-tap kernel on the fold's quantised taps, so a
            #  sample's result is bit-identical whatever the size of the batch it arrives in)
            fold = fold_w and B * T >= self.fold_min_frames
            if fold:
                self.flop_scale[s.name] = ops.fold_plan(g.Hin, g.Win, g.Hout, g.Wout, B, str(dev)).taps_per_output / 25.0
            else:
                self.flop_scale.pop(s.name, None)
            # the fp32 [K][Cout] copy is only read by the CUDA-core kernels (forward impl='simt', backward bwd_impl='simt')
            w_kn, w_i8 = s.packed(self.weight_planes, use_i8, need_kn=(want_h and self.bwd_impl == 'simt') or not use_i8, fold=fold_w)
            if fold_w and not fold:
                w_i8 = (s.unfolded_fold_taps(self.weight_planes, w_i8[4]), w_i8[3])     # (25-tap image of the fold's taps, wscale)
            decay = params[2 * i + 1]
            if decay is not None:
                decay = decay.detach().contiguous()
            node = s.node
            v_in = _node_v_in(node, (B, g.Hout, g.Wout, g.Cout), dev)
            resid = acts[s.resid] if s.resid is not None else None
            if self.timing is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            common = dict(T=T, B=B, neuron=node.kind, gain=s.gain_mod.gain(), v_th=node.v_threshold, v_reset=node.v_reset,
                          tau=node._tau_value(), decay=decay, v_in=v_in, want_v_out=self.keep_state, resid=resid, want_h=want_h)
            # per block: the [k, B / k] view is the same memory as [1, B], so only the blocks that stream many weight channel
            # blocks per tile take it (the full-resolution blocks have nothing to amortise and lose tile-level parallelism)
            k_steps = k_all if g.Cin >= (self.bas_min_cin if B // k_all < 8 else min(self.bas_min_cin, 64)) else 1
            as_steps = (lambda a, k=k_steps: a.view(k, B // k, *a.shape[2:])) if k_steps > 1 else (lambda a: a)
            if k_steps > 1:
                common.update(T=k_steps, B=B // k_steps, independent_steps=True, resid=as_steps(resid) if resid is not None else None)
            if use_i8:
                if first and not packed_in:
                    status = self._status_word(dev) if self.check_input else None
                    xin = ops.pack_events(x_seq, status)     # fp32 NCHW counts -> u8 NHWC4
                    acts['x_packed'] = xin
                    if status is not None and not torch.cuda.is_current_stream_capturing():
                        self._record_input_status()
                tsum = None
                if s.out in head_srcs and self.heads_time_sum and 1 < T <= 86:
                    tsum = torch.empty((B, g.Hout, g.Wout, g.Cout), dtype=ops.ACT_DTYPE, device=dev)
                    tsums[s.out] = tsum
                if fold:
                    out, v_out, h_seq = ops.conv_i8_fwd_folded(as_steps(xin), g, w_i8[0], w_i8[1], w_i8[2], w_i8[3],
                                                               planes=self.weight_planes, tsum=tsum,
                                                               stats=stats[i] if stats is not None else None, **common)
                else:
                    out, v_out, h_seq = ops.conv_i8_fwd(as_steps(xin), g, w_i8[0], w_i8[1], planes=self.weight_planes,
                                                        cin=ops.first_layer_channels(g.Cin) if first else g.Cin, tsum=tsum,
                                                        stats=stats[i] if stats is not None else None, **common)
                if k_steps > 1:
                    out = out.view(T, B, *out.shape[2:])
            else:
                out, v_out, h_seq = ops.conv_neuron_fwd(xin, g, w_kn, in_layout=SS_IN_F32_BTCHW if first else SS_IN_U8_TBHWC,
                                                        **common)
            if self.timing is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                self.timing.append((s.name, ev0, ev1))
            acts[s.out] = out
            if self.keep_state:
                node.v = v_out.permute(0, 3, 1, 2)     # NCHW-shaped view, as the reference exposes it
                node._v_from_grad_call = bool(want_h)
            saved['sites'].append({'geom': g, 'w_kn': w_kn, 'decay': decay, 'h_seq': h_seq, 'v_in': v_in})
        if not self.heads:
            # stand-alone block (run_sew_block): no readout
            saved['depths'] = None
            side['acts'] = acts
            if not want_h:
                saved['acts'] = None
            return saved
        # heads + I-neurons
        hg, hw, hb, hacts = [], [], [], []
        for j, h in enumerate(self.heads):
            a = acts[h.src]
            g = h.geom(int(a.shape[2]), int(a.shape[3]))
            hg.append(g)
            w = params[n_site + 2 * j].detach()
            # [9][C] tap-major copy of the head weight, cached on the parameter's version (one small copy kernel per head and call)
            key = (w.data_ptr(), params[n_site + 2 * j]._version, str(w.device), g.Cin)
            if getattr(h, '_w9c', None) is None or h._w9c[0] != key:
                h._w9c = (key, w[0].permute(1, 2, 0).reshape(9, g.Cin).contiguous().float())
            hw.append(h._w9c[1])
            hb.append(params[n_site + 2 * j + 1].detach().reshape(1).contiguous().float())
            hacts.append(a)
        H, W = hg[0].Hout, hg[0].Wout
        vi = self.ineuron.v
        if isinstance(vi, torch.Tensor):
            v_io = vi.detach().to(device=dev, dtype=torch.float32).reshape(B, H, W).contiguous().clone()
        else:
            v_io = torch.full((B, H, W), float(vi), dtype=torch.float32, device=dev)
        gain = self.heads[0].gain_mod.gain()
        if self.timing is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        acts_sum = [tsums[h.src] for h in self.heads] if len(tsums) == len(self.heads) else None
        depths = ops.heads_fwd(hacts, hg, hw, hb, T=T, B=B, H=H, W=W, gain=gain, v_io=v_io, acts_sum=acts_sum)
        if self.timing is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self.timing.append(('heads', ev0, ev1))
        self.ineuron.v = v_io.view(B, 1, H, W)
        saved.update({'hg': hg, 'hw': hw, 'hacts': hacts, 'gain': gain, 'H': H, 'W': W, 'depths': depths})
        side['acts'] = acts
        if not want_h:
            saved['acts'] = None
        return saved

    # ------------------------------------------------------------------ backward
    def _run_backward(self, saved, params, n_site, g_depths, inject=None, input_grad_of=None):
        """Returns the parameter gradients (and, for stand-alone blocks, the gradient buffer of activation ``input_grad_of``)."""
        L = _lib.lib()
        B, T = saved['B'], saved['T']
        H, W = saved.get('H'), saved.get('W')
        acts = saved['acts']
        dev = g_depths.device if g_depths is not None else next(iter(inject.values())).device
        grads = [None] * len(params)
        g = {}
        hook = self.grad_hook
        if hook is not None:
            hook.begin()

        def gbuf(name, fresh_ok=False):
            if name not in g:
                alloc = torch.empty if fresh_ok else torch.zeros      # fresh_ok: the caller writes every element
                g[name] = alloc(acts[name].shape, dtype=torch.float32, device=dev)
            return g[name]

        # ---- heads
        if self.heads:
            a = _lib.HeadsArgs()
            a.T, a.B, a.H, a.W, a.gain = T, B, H, W, saved['gain']
            keep = []
            vp4 = ctypes.c_void_p * 4
            g_acts, g_w, g_b, bins = vp4(), vp4(), vp4(), vp4()
            gw_t, gb_t = [], []
            store_heads = len({h.src for h in self.heads}) == len(self.heads)     # distinct sources: each buffer has one writer here
            for j, h in enumerate(self.heads):
                gm = saved['hg'][j]
                ym, xm = gm.maps(dev)
                a.C[j], a.Hs[j], a.Ws[j] = gm.Cin, gm.Hin, gm.Win
                a.acts[j] = saved['hacts'][j].data_ptr()
                a.w[j] = saved['hw'][j].data_ptr()
                a.ymap[j], a.xmap[j] = ym.data_ptr(), xm.data_ptr()
                ga = gbuf(h.src, fresh_ok=store_heads)
                gw = torch.zeros((9, gm.Cin), dtype=torch.float32, device=dev)
                gb = torch.zeros((1,), dtype=torch.float32, device=dev)
                bn = torch.zeros((2, B, gm.Hin, gm.Win, 9), dtype=torch.float32, device=dev)
                keep += [ym, xm, bn]
                gw_t.append(gw)
                gb_t.append(gb)
                g_acts[j], g_w[j], g_b[j], bins[j] = ga.data_ptr(), gw.data_ptr(), gb.data_ptr(), bn.data_ptr()
            _lib.check(L.ss_heads_bwd(ctypes.byref(a), _ptr(g_depths), g_acts, g_w, g_b, bins, 1 if store_heads else 0, _stream()),
                       'ss_heads_bwd')
            for j, h in enumerate(self.heads):
                C = saved['hg'][j].Cin
                grads[n_site + 2 * j] = gw_t[j].reshape(3, 3, C).permute(2, 0, 1).reshape(1, C, 3, 3).contiguous()
                grads[n_site + 2 * j + 1] = gb_t[j]
                if hook is not None:
                    hook.ready(grads[n_site + 2 * j])
                    hook.ready(grads[n_site + 2 * j + 1])

        # ---- gradients arriving on the returned spike maps (last timestep, NCHW) join the buffers the heads just created
        for name, gs in (inject or {}).items():
            buf = gbuf(name)
            buf[T - 1] += gs.permute(0, 2, 3, 1).to(torch.float32)

        # ---- spiking blocks, reverse order
        main = torch.cuda.current_stream(dev)
        side = None
        if self.wgrad_stream and self.bwd_impl == 'umma' and not torch.cuda.is_current_stream_capturing():
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
        side_keep = []
        for i in range(len(self.sites) - 1, -1, -1):
            s, sv = self.sites[i], saved['sites'][i]
            gm = sv['geom']
            if s.out not in g:
                continue        # nothing downstream asked for a gradient
            g_out = g.pop(s.out)
            node = s.node
            N = B * gm.Hout * gm.Wout * gm.Cout
            first = s.src == 'x'
            x_in = acts['x_packed'] if (first and 'x_packed' in acts) else acts[s.src]
            tc_ok = self.bwd_impl == 'umma' and gm.Cout % 16 == 0 and x_in.dtype == ops.ACT_DTYPE
            tc_w = tc_ok and (x_in.shape[-1] % 16 == 0 or (first and x_in.shape[-1] == 4 and gm.ks == 5 and gm.stride == 1))
            tc_d = tc_ok and not first and gm.Cin % 32 == 0
            need32 = not (tc_w and (tc_d or first))
            g_acc = torch.empty_like(g_out) if need32 else None
            g_b16 = torch.empty(g_out.shape, dtype=torch.bfloat16, device=dev) if (tc_w or tc_d) else None
            decay = sv['decay']
            g_decay = torch.zeros((1,), dtype=torch.float32, device=dev) if decay is not None else None
            sf = node.surrogate_function
            rc = L.ss_neuron_bwd_ex(T, N, node.kind, sf.kind, sf.alpha, s.gain_mod.gain(), node.v_threshold, node.v_reset,
                                    node._tau_value(), _ptr(decay), _ptr(sv['h_seq']), _ptr(sv['v_in']), _ptr(g_out), None,
                                    _ptr(g_acc), _ptr(g_b16), None, _ptr(g_decay), _stream())
            _lib.check(rc, 'ss_neuron_bwd')
            if s.resid is not None:
                if s.resid in g:
                    g[s.resid] += g_out
                else:
                    g[s.resid] = g_out      # donate: the residual branch passes the gradient through unchanged
            cg = _lib.ConvGeom(T=T, B=B, Hin=gm.Hin, Win=gm.Win, Cin=gm.Cin, Hout=gm.Hout, Wout=gm.Wout, Cout=gm.Cout,
                               ks=gm.ks, in_layout=SS_IN_F32_BTCHW if first else SS_IN_U8_TBHWC, neuron=node.kind,
                               reserved0=0, gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0, reserved1=0, reserved2=0)
            ym, xm = gm.maps(dev)
            if tc_w and side is not None:
                # The weight gradient hangs off the critical path (scan -> data gradient -> next block's scan): it runs on a second
                # stream, so that its CTAs and the data-gradient / scan kernels of the following blocks fill each other's partly empty
                # rounds (a weight-gradient grid is one wave of 128-148 CTAs of unequal length; both kinds of CTA take a whole SM).
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    cin_dev = int(x_in.shape[-1])
                    g_wkn = ops.conv_wgrad_bf16(x_in, g_b16, gm, T, B, cin=cin_dev, x_full_range=first)
                    if cin_dev != gm.Cin:       # packed first layer: drop the padding channels
                        g_wkn = g_wkn.view(gm.ks * gm.ks, cin_dev, gm.Cout)[:, :gm.Cin].reshape(gm.K, gm.Cout)
                    grads[2 * i] = ops.kn_to_weight(g_wkn, gm.Cout, gm.Cin, gm.ks)
                    if hook is not None:
                        hook.ready(grads[2 * i])
                # (no record_stream: it defers the allocator's reuse of these blocks and the step time then takes ~10 steps to settle;
                #  the operands are simply kept alive until the main stream has waited for the side stream below, and the gradient's
                #  block is only ever re-used by the side stream, which is ordered after everything the main stream did with it)
                side_keep.append((x_in, g_b16))
            else:
                if tc_w:
                    cin_dev = int(x_in.shape[-1])
                    g_wkn = ops.conv_wgrad_bf16(x_in, g_b16, gm, T, B, cin=cin_dev, x_full_range=first)   # event counts may exceed 127
                    if cin_dev != gm.Cin:       # packed first layer: drop the padding channels
                        g_wkn = g_wkn.view(gm.ks * gm.ks, cin_dev, gm.Cout)[:, :gm.Cin].reshape(gm.K, gm.Cout)
                else:
                    g_wkn = torch.zeros((gm.K, gm.Cout), dtype=torch.float32, device=dev)
                    rc = L.ss_conv_wgrad(ctypes.byref(cg), _ptr(acts[s.src]), _ptr(ym), _ptr(xm), _ptr(g_acc), _ptr(g_wkn), _stream())
                    _lib.check(rc, 'ss_conv_wgrad')
                grads[2 * i] = ops.kn_to_weight(g_wkn, gm.Cout, gm.Cin, gm.ks)
                if hook is not None:
                    # all-reduce this block's gradient now, on NCCL's stream, while the earlier layers' gradients are computed
                    hook.ready(grads[2 * i])
            if g_decay is not None:
                grads[2 * i + 1] = g_decay.reshape(params[2 * i + 1].shape)
            if hook is not None:
                hook.ready(grads[2 * i + 1])
            if not first:
                gx = gbuf(s.src)
                if tc_d:
                    s.dgrad_plan(gm).run(g_b16, gx, T, B)
                else:
                    w_kn = sv['w_kn'] if sv['w_kn'] is not None else ops.weight_to_kn(s.conv.weight)
                    rc = L.ss_conv_dgrad(ctypes.byref(cg), _ptr(ym), _ptr(xm), _ptr(w_kn), _ptr(g_acc), _ptr(gx), _stream())
                    _lib.check(rc, 'ss_conv_dgrad')
            del g_acc, g_b16, g_out
        if side is not None:
            main.wait_stream(side)          # every weight gradient is complete before autograd hands them on
            side_keep.clear()
        if hook is not None:
            hook.finish()
        if input_grad_of is not None:
            return grads, g.get(input_grad_of)
        return grads


# ---------------------------------------------------------------------------------------- stand-alone blocks
def _nchw_to_tbhwc(x):
    """[B,C,H,W] fp32 spikes -> u8 [1,B,H,W,C] (layout plumbing for stand-alone block calls)."""
    ops._require_cuda(x, 'x')
    return x.detach().permute(0, 2, 3, 1).to(ops.ACT_DTYPE).contiguous().unsqueeze(0)


def _tbhwc_to_nchw(a):
    return a[0].permute(0, 3, 1, 2).float()


class _BlockFunction(torch.autograd.Function):
    """Stand-alone spiking blocks (SEWResBlock.forward on its own): a head-less engine over the block's sites.  Input / output are
    the reference's NCHW fp32 spike tensors, one timestep per call; the gradient flows to the block input and to the weights
    through the same kernels as inside the models (surrogate scan, tensor-core dgrad / wgrad)."""

    @staticmethod
    def forward(ctx, eng, x, in_name, out_name, need_grad, *params):
        xb = _nchw_to_tbhwc(x)
        side = {'need_grad': need_grad, 'return_layers': False, 'spike_outputs': (), 'seed_acts': {in_name: xb}}
        n_site = len(params)
        res = eng._run_forward(xb, params, n_site, side, want_h=need_grad)
        out = _tbhwc_to_nchw(side['acts'][out_name])
        ctx.saved = None
        if need_grad:
            ctx.eng, ctx.params, ctx.n_site, ctx.saved = eng, params, n_site, res
            ctx.in_name, ctx.out_name = in_name, out_name
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.saved is None:
            raise RuntimeError('stereospike_b200: backward through the same forward twice (the saved potentials were released)')
        grads, g_in = ctx.eng._run_backward(ctx.saved, ctx.params, ctx.n_site, None, {ctx.out_name: g_out.contiguous().float()},
                                            input_grad_of=ctx.in_name)
        ctx.saved = None
        gx = g_in[0].permute(0, 3, 1, 2).contiguous() if g_in is not None else None
        return (None, gx, None, None, None) + tuple(grads)


def run_sew_block(blk, x):
    """SEWResBlock.forward for a stand-alone call: x [B,C,H,W] fp32 spikes -> [B,C,H,W] fp32, stateful (one timestep per call),
    differentiable w.r.t. x and the block's parameters (reference network/blocks.py:161-171 under autograd)."""
    ops._require_cuda(x, 'x')
    if not hasattr(blk, '_engine'):
        sites = [Site('conv1', 'mid', 'in', blk.conv1[0], blk.conv1[1], blk.sn1),
                 Site('conv2', 'out', 'mid', blk.conv2[0], blk.conv2[1], blk.sn2, resid='in')]
        eng = Engine(sites, [], None)
        eng.check_input = False
        object.__setattr__(blk, '_engine', eng)
    eng = blk._engine
    if blk.conv1[0].in_channels % 32 != 0:
        raise NotImplementedError('stand-alone SEWResBlock: the fused kernels tile 32 channels (every SEWResBlock of the reference '
                                  f'models has 512); got {blk.conv1[0].in_channels}')
    params, _ = eng._flat_params()
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p is not None and p.requires_grad for p in params))
    if need_grad:
        eng._warn_truncated_bptt()
    return _BlockFunction.apply(eng, x, 'in', 'out', need_grad, *params)


class _LinearBlockFunction(torch.autograd.Function):
    """NNConvUpsampling on its own, any channel counts / kernel size, arbitrary fp32 input: upsample-gather + valid conv on the fp32
    CUDA-core kernel (a non-firing IF step from rest returns the conv result as its potential), gradients through ss_conv_dgrad /
    ss_conv_wgrad.  The 1-channel 3x3 heads inside the models go through the dedicated heads kernels instead."""

    @staticmethod
    def forward(ctx, x, weight, bias, up_size, conv=None):
        B, C, Hs, Ws = x.shape
        co, ci, ks, _ = weight.shape
        cp = (co + 31) // 32 * 32                               # the kernel tiles 32 output channels
        if conv is None:
            g = BlockGeom('upconv', C, cp, ks, Hs, Ws, up_size[0], up_size[1])
        else:                                                   # plain Conv2d (stride, padding): the analog comparison model
            stride, pad = conv
            g = BlockGeom('conv', C, cp, ks, Hs, Ws, conv_out_size(Hs, ks, stride, pad), conv_out_size(Ws, ks, stride, pad),
                          stride, pad)
        w_kn = torch.zeros((ks * ks * C, cp), dtype=torch.float32, device=x.device)
        w_kn[:, :co] = ops.weight_to_kn(weight.detach().float())
        x5 = x.detach().float().contiguous().view(B, 1, C, Hs, Ws)
        _, _, h = ops.conv_neuron_fwd(x5, g, w_kn, T=1, B=B, in_layout=SS_IN_F32_BTCHW, neuron=_lib.SS_NEURON_IF, gain=1.0,
                                      v_th=3.0e38, v_reset=0.0, want_h=True)
        y = h[0, :, :, :, :co].permute(0, 3, 1, 2)
        if bias is not None:
            y = y + bias.detach().float().view(1, co, 1, 1)
        ctx.save_for_backward(x5, w_kn)
        ctx.geom, ctx.co = g, co
        return y.contiguous()

    @staticmethod
    def backward(ctx, g_y):
        x5, w_kn = ctx.saved_tensors
        g, co = ctx.geom, ctx.co
        B = int(x5.shape[0])
        dev = g_y.device
        L = _lib.lib()
        g_acc = torch.zeros((1, B, g.Hout, g.Wout, g.Cout), dtype=torch.float32, device=dev)
        g_acc[0, :, :, :, :co] = g_y.float().permute(0, 2, 3, 1)
        cg = _lib.ConvGeom(T=1, B=B, Hin=g.Hin, Win=g.Win, Cin=g.Cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                           in_layout=SS_IN_F32_BTCHW, neuron=_lib.SS_NEURON_IF, reserved0=0, gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0,
                           reserved1=0, reserved2=0)
        ym, xm = g.maps(dev)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            g_x = torch.zeros((1, B, g.Hin, g.Win, g.Cin), dtype=torch.float32, device=dev)
            _lib.check(L.ss_conv_dgrad(ctypes.byref(cg), _ptr(ym), _ptr(xm), _ptr(w_kn), _ptr(g_acc), _ptr(g_x), _stream()), 'ss_conv_dgrad')
            gx = g_x[0].permute(0, 3, 1, 2).contiguous()
        if ctx.needs_input_grad[1]:
            g_wkn = torch.zeros((g.K, g.Cout), dtype=torch.float32, device=dev)
            _lib.check(L.ss_conv_wgrad(ctypes.byref(cg), _ptr(x5), _ptr(ym), _ptr(xm), _ptr(g_acc), _ptr(g_wkn), _stream()), 'ss_conv_wgrad')
            gw = ops.kn_to_weight(g_wkn, g.Cout, g.Cin, g.ks)[:co].contiguous()
        if ctx.needs_input_grad[2]:
            gb = g_y.float().sum(dim=(0, 2, 3))
        return gx, gw, gb, None, None


def run_dense_conv(conv, x):
    """nn.Conv2d.forward (square kernel, one stride / padding for both axes, groups = dilation = 1) on arbitrary fp32 input through
    the fp32 CUDA-core kernel, with gradients: the convolutions of the analog comparison model (network/ANN_models.py:39-72)."""
    ops._require_cuda(x, 'x')
    ks, st, pd = conv.kernel_size, conv.stride, conv.padding
    if ks[0] != ks[1] or st[0] != st[1] or pd[0] != pd[1] or conv.groups != 1 or tuple(conv.dilation) != (1, 1) or \
            conv.padding_mode != 'zeros':
        raise NotImplementedError('run_dense_conv: square kernels, equal strides / zero paddings, groups = dilation = 1')
    return _LinearBlockFunction.apply(x, conv.weight, conv.bias, None, (int(st[0]), int(pd[0])))


def run_linear_block(up, x):
    """NNConvUpsampling.forward for a stand-alone call (no neuron; reference network/blocks.py:130-132).  The 1-channel 3x3 heads
    (the only stand-alone use in the reference, predict_depthK) run on the heads kernels when the input is a uint8 spike-count
    tensor and no gradient is needed; every other call -- any fp32 input, whose values may be arbitrary reals as in the analog
    comparison model -- runs the general fp32 kernel with gradients w.r.t. input, weight and bias."""
    conv = up.up[1]
    ops._require_cuda(x, 'x')
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in conv.parameters()))
    spikes_like = x.dtype == torch.uint8 and conv.in_channels % 16 == 0
    if need_grad or conv.out_channels != 1 or conv.kernel_size[0] != 3 or not spikes_like:
        return _LinearBlockFunction.apply(x, conv.weight, conv.bias, up.up_size, None)
    B, C, Hs, Ws = x.shape
    xb = _nchw_to_tbhwc(x)
    g = BlockGeom('upconv', C, 1, 3, Hs, Ws, up.up_size[0], up.up_size[1])
    dev = x.device
    zero_act = [xb, xb, xb, xb]
    zw = torch.zeros((9, C), dtype=torch.float32, device=dev)
    zb = torch.zeros((1,), dtype=torch.float32, device=dev)
    w = conv.weight.detach()[0].permute(1, 2, 0).reshape(9, C).contiguous().float()
    b = conv.bias.detach().reshape(1).float() if conv.bias is not None else zb
    v = torch.zeros((B, g.Hout, g.Wout), dtype=torch.float32, device=dev)
    d = ops.heads_fwd(zero_act, [g] * 4, [w, zw, zw, zw], [b, zb, zb, zb], T=1, B=B, H=g.Hout, W=g.Wout, gain=1.0, v_io=v)
    return d[0].unsqueeze(1)
