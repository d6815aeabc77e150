"""Drop-in replacements for the spiking U-Nets of the reference's ``network/SNN_models.py``: same class names,
constructor signatures, sub-module nesting (hence state-dict keys: ``bottom.0.weight``, ``conv1.0.weight``,
``bottleneck.0.conv1.0.weight``, ``deconv4.0.up.1.weight``, ``predict_depth4.0.up.1.{weight,bias}``, PLIF
``*.2.w`` / ``bottleneck.N.snK.w``) and the same helper methods.

    NeuromorphicNet                                                   SNN_models.py:11-60
    StereoSpike                                                       SNN_models.py:63-248   (IF, binocular)
    fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike       SNN_models.py:251-435  (LIF / PLIF, binocular)
    fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike  SNN_models.py:438-622  (LIF / PLIF, Cin = 2)

``forward(x)`` keeps the reference contract: ``x`` is ``[B, >=1, C, 260, 346]`` and only frame 0 is read; neuron
state persists between calls until ``functional.reset_net``.  ``forward_seq(x_seq)`` is the opt-in T-loop
(``[B, T, C, H, W]``, equivalent to calling ``forward`` once per timestep without reset) executed as ONE fused
kernel per block with the membrane potential held on-chip across the T loop.
"""
import torch
import torch.nn as nn

from . import neuron, surrogate
from .blocks import MultiplyBy, NNConvUpsampling, SEWResBlock
from .engine import Engine, Head, Site

ENCODER = (('conv1', 32, 64), ('conv2', 64, 128), ('conv3', 128, 256), ('conv4', 256, 512))
DECODER = (('deconv4', 512, 256, (33, 44)), ('deconv3', 256, 128, (65, 87)),
           ('deconv2', 128, 64, (130, 173)), ('deconv1', 64, 32, (260, 346)))
HEADS = (('predict_depth4', 256), ('predict_depth3', 128), ('predict_depth2', 64), ('predict_depth1', 32))
LAYER_NAMES = ('out_bottom', 'out_conv1', 'out_conv2', 'out_conv3', 'out_conv4', 'out_rconv', 'out_deconv4',
               'out_add4', 'out_deconv3', 'out_add3', 'out_deconv2', 'out_add2', 'out_deconv1', 'out_add1')


class SpikeList(list):
    """The list of spike maps a model returns; ``penalty`` (when gradients are enabled and fp32 maps were asked for) is the
    reference's SpikePenalization_Loss of these maps as a differentiable scalar, computed from counters the block epilogues
    accumulate -- stereospike_b200.loss.Total_Loss(penalize_spikes=True) uses it instead of a pass over the maps."""
    penalty = None


class NeuromorphicNet(nn.Module):
    def __init__(self, surrogate_function=None, detach_reset=True, v_threshold=1.0, v_reset=0.0):
        super().__init__()
        self.surrogate_fct = surrogate_function if surrogate_function is not None else surrogate.Sigmoid()
        self.detach_rst = detach_reset
        self.v_th = v_threshold
        self.v_rst = v_reset

        self.max_test_accuracy = float('inf')
        self.epoch = 0

    # ---- state helpers (SNN_models.py:22-60)
    def detach(self):
        for m in self.modules():
            if isinstance(m, neuron.BaseNode) and isinstance(m.v, torch.Tensor):
                m.v = m.v.detach()
                m._v_from_grad_call = False        # an explicit truncation point, like in the reference (train.py:242)

    def get_network_state(self):
        return [m.v for m in self.modules() if hasattr(m, 'reset')]

    def change_network_state(self, new_state):
        module_index = 0
        for m in self.modules():
            if hasattr(m, 'reset'):
                m.v = new_state[module_index]
                module_index += 1

    def set_output_potentials(self, new_pots):
        module_index = 0
        for m in self.modules():
            if isinstance(m, neuron.IFNode):
                m.v = new_pots[module_index]
                module_index += 1

    def increment_epoch(self):
        self.epoch += 1

    def get_max_accuracy(self):
        return self.max_test_accuracy

    def update_max_accuracy(self, new_acc):
        self.max_test_accuracy = new_acc

    def count_trainable_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class _SpikingUNet(NeuromorphicNet):
    """Shared wiring of the three reference classes (they are copies of one U-Net upstream)."""

    _returns_spikes = True

    def _build(self, cin0, make_outer, make_inner_kwargs, multiply_factor, ineuron):
        g = multiply_factor
        self.bottom = nn.Sequential(nn.Conv2d(cin0, 32, kernel_size=5, stride=1, padding=2, bias=False),
                                    MultiplyBy(g), make_outer())
        for name, ci, co in ENCODER:
            setattr(self, name, nn.Sequential(nn.Conv2d(ci, co, kernel_size=5, stride=2, padding=2, bias=False),
                                              MultiplyBy(g), make_outer()))
        self.bottleneck = nn.Sequential(SEWResBlock(512, connect_function='ADD', multiply_factor=g, **make_inner_kwargs),
                                        SEWResBlock(512, connect_function='ADD', multiply_factor=g, **make_inner_kwargs))
        for name, ci, co, up in DECODER:
            setattr(self, name, nn.Sequential(NNConvUpsampling(ci, co, kernel_size=5, up_size=up), MultiplyBy(g),
                                              make_outer()))
        for name, ci in HEADS:
            setattr(self, name, nn.Sequential(NNConvUpsampling(ci, 1, kernel_size=3, up_size=(260, 346), bias=True),
                                              MultiplyBy(g)))
        self.Ineurons = ineuron
        object.__setattr__(self, '_engine', None)

    # ---- engine
    @property
    def engine(self):
        if self._engine is None:
            sites = [Site('bottom', 'out_bottom', 'x', self.bottom[0], self.bottom[1], self.bottom[2])]
            prev = 'out_bottom'
            for name, _, _ in ENCODER:
                seq = getattr(self, name)
                sites.append(Site(name, 'out_' + name, prev, seq[0], seq[1], seq[2]))
                prev = 'out_' + name
            for bi, blk in enumerate(self.bottleneck):
                mid, out = f'_sew{bi}_mid', ('out_rconv' if bi == len(self.bottleneck) - 1 else f'_sew{bi}_out')
                sites.append(Site(f'bottleneck.{bi}.conv1', mid, prev, blk.conv1[0], blk.conv1[1], blk.sn1))
                sites.append(Site(f'bottleneck.{bi}.conv2', out, mid, blk.conv2[0], blk.conv2[1], blk.sn2, resid=prev))
                prev = out
            skips = ('out_conv3', 'out_conv2', 'out_conv1', 'out_bottom')
            heads = []
            for (name, _, _, up), skip, (hname, _) in zip(DECODER, skips, HEADS):
                seq = getattr(self, name)
                out = 'out_add' + name[-1]
                sites.append(Site(name, out, prev, seq[0].up[1], seq[1], seq[2], resid=skip, up_size=up))
                hseq = getattr(self, hname)
                heads.append(Head(hname, out, hseq[0], hseq[1]))
                prev = out
            object.__setattr__(self, '_engine', Engine(sites, heads, self.Ineurons))
        return self._engine

    def set_kernel_options(self, impl=None, weight_planes=None, keep_state=None, heads_time_sum=None, fold_upsample=None,
                           bwd_impl=None, fold_min_frames=None, batch_as_steps=None):
        """impl: 'umma' (tcgen05 int8 tensor-core kernel; default) or 'simt' (exact-fp32 CUDA cores).
        weight_planes: int8 digit planes per weight -- 3 = 24-bit fixed point, fp32-class (default);
        2 = 16-bit (the reduced-precision training configuration); 4 = 32-bit.
        bwd_impl: 'umma' (conv gradients on the bf16 tensor cores, fp32 accumulation; default) or 'simt' (fp32 CUDA cores).
        fold_upsample / fold_min_frames: NNConvUpsampling blocks as folded 3x3 convs (default on) for calls of at least
        fold_min_frames event frames (B * T, default 16; smaller calls are launch-latency-bound and keep the single 25-tap launch).
        batch_as_steps: stateless single-step calls on a batch run as k independent steps of B / k samples per launch (default on)."""
        e = self.engine
        if impl is not None:
            assert impl in ('auto', 'umma', 'simt')
            e.impl = impl
        if weight_planes is not None:
            assert weight_planes in (2, 3, 4)
            e.weight_planes = weight_planes
        if keep_state is not None:
            e.keep_state = bool(keep_state)
        if heads_time_sum is not None:
            e.heads_time_sum = bool(heads_time_sum)
        if fold_upsample is not None:       # True / False, or the names of the decoder blocks to fold, e.g. ('deconv4', 'deconv3')
            e.fold_upsample = frozenset(fold_upsample) if isinstance(fold_upsample, (set, frozenset, list, tuple)) else bool(fold_upsample)
        if bwd_impl is not None:
            assert bwd_impl in ('umma', 'simt')
            e.bwd_impl = bwd_impl
        if fold_min_frames is not None:
            e.fold_min_frames = int(fold_min_frames)
        if batch_as_steps is not None:
            e.batch_as_steps = bool(batch_as_steps)
        return self

    # ---- forward
    _SPIKE_OUTPUTS = ('out_rconv', 'out_add4', 'out_add3', 'out_add2', 'out_add1')

    def _package(self, depths, side, spikes_fp32):
        d = [depths[3].unsqueeze(1), depths[2].unsqueeze(1), depths[1].unsqueeze(1), depths[0].unsqueeze(1)]
        if not self._returns_spikes:
            return d
        if spikes_fp32 and len(side.get('spikes_fp32', ())) >= len(self._SPIKE_OUTPUTS):
            # differentiable: a loss on them reaches the weights through the surrogates
            out = SpikeList(side['spikes_fp32'][:len(self._SPIKE_OUTPUTS)])
            if len(side['spikes_fp32']) > len(self._SPIKE_OUTPUTS):
                out.penalty = side['spikes_fp32'][len(self._SPIKE_OUTPUTS)]
            return d, out
        acts = side['acts']
        spks = []
        for k in self._SPIKE_OUTPUTS:
            s = acts[k][-1].permute(0, 3, 1, 2)        # last timestep, NCHW-shaped view of the u8 NHWC buffer
            spks.append(s.float() if spikes_fp32 else s)
        return d, spks

    def _run(self, x, spikes_fp32):
        want = self._SPIKE_OUTPUTS if (spikes_fp32 and self._returns_spikes) else None
        depths, side = self.engine.run(x, spike_outputs=want)
        return self._package(depths, side, spikes_fp32)

    def forward(self, x):
        """Reference contract (SNN_models.py:152-192): reads frame 0 of ``x``; stateful across calls."""
        return self._run(x[:, 0:1], True)

    def forward_seq(self, x_seq, spikes_fp32=False):
        """T-loop over ``x_seq[:, t]`` without reset, fused: one kernel per block for all T timesteps.
        ``x_seq`` is the reference's fp32 ``[B, T, C, H, W]`` or the packed u8 ``[T, B, H, W, 4]`` of ``stereospike_b200.events``.
        Returns what the LAST ``forward`` call of the equivalent loop would return.  Spike tensors are
        NCHW-shaped u8 views unless ``spikes_fp32`` (then fp32, and differentiable when gradients are enabled)."""
        return self._run(x_seq, spikes_fp32)

    def set_init_depths_potentials(self, depth_prior):
        self.Ineurons.v = depth_prior

    def calculate_firing_rates(self, x):
        """Per-layer spike densities count_nonzero/numel of frame 0 (SNN_models.py:194-245).  Like the reference, this
        advances the neuron state by one step.  The counts come from the block epilogues (side['stats']: spikes fired and
        nonzero outputs per block, accumulated with dp4a from the registers that hold the spikes) -- one small device-to-host
        copy instead of 14 count_nonzero passes over the activations."""
        eng = self.engine
        keep = eng.collect_stats
        eng.collect_stats = True
        try:
            with torch.no_grad():
                _, side = eng.run(x[:, 0:1])
        finally:
            eng.collect_stats = keep
        acts = side['acts']
        if side.get('stats') is None:           # CUDA-core path (impl='simt'): count on the stored activations
            rates = {}
            for k in LAYER_NAMES:
                if k.startswith('out_deconv'):
                    n = k[-1]
                    skip = {'4': 'out_conv3', '3': 'out_conv2', '2': 'out_conv1', '1': 'out_bottom'}[n]
                    t = acts['out_add' + n][-1].float() - acts[skip][-1].float()
                else:
                    t = acts[k][-1]
                rates[k] = float(t.count_nonzero()) / t.numel()
            return rates
        st = side['stats'].cpu()
        rates = {}
        for k in LAYER_NAMES:
            fired = k.startswith('out_deconv')                  # the block's own spikes, before the skip connection is added
            name = 'out_add' + k[-1] if fired else k
            i = eng.site_index_of_output(name)
            rates[k] = float(st[i, 0 if fired else 1]) / acts[name][-1].numel()
        return rates


class StereoSpike(_SpikingUNet):
    """Baseline binocular model: IF neurons, ATan surrogate outside the bottleneck, default Sigmoid inside
    (SNN_models.py:63-150; the reference does not forward v_threshold / v_reset to its base, so they are
    always 1.0 / 0.0 -- reproduced here).

    ``in_channels`` (extension, all three classes): the channel-concatenated temporal mode of the reference's scripts
    (train.py:206-218: ``nfpdm`` frames per depth map folded into the channel axis, "number of filters in the first convolution
    should be changed accordingly") -- e.g. ``in_channels = 2 * nfpdm * 2`` for the binocular model.  More than 4 channels run the
    first block as an ordinary 32-channel tensor-core block on frames packed to u8 [T,B,H,W,32]."""

    def __init__(self, surrogate_function=None, detach_reset=True, v_threshold=1.0, v_reset=0.0, multiply_factor=1.,
                 in_channels=4):
        super().__init__(surrogate_function=surrogate_function, detach_reset=detach_reset)
        sf = self.surrogate_fct
        outer = lambda: neuron.IFNode(v_threshold=self.v_th, v_reset=self.v_rst, surrogate_function=sf, detach_reset=True)
        self._build(in_channels, outer, dict(v_threshold=self.v_th, v_reset=self.v_rst), multiply_factor,
                    neuron.IFNode(v_threshold=float('inf'), v_reset=0.0, surrogate_function=sf))


class fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(_SpikingUNet):
    """LIF (``use_plif=False``: LIFNode + ATan) or PLIF (default-Sigmoid surrogate) binocular model; the
    bottleneck is always PLIF + Sigmoid (SNN_models.py:251-340)."""

    _cin0 = 4

    def __init__(self, use_plif=False, detach_reset=True, tau=10., v_threshold=1.0, v_reset=0.0, multiply_factor=1.,
                 in_channels=None):
        super().__init__(detach_reset=detach_reset)
        self.is_cext_model = False
        if in_channels is not None:
            self._cin0 = int(in_channels)
        if use_plif:
            outer = lambda: neuron.ParametricLIFNode(init_tau=tau, v_threshold=v_threshold, v_reset=v_reset,
                                                     detach_reset=True)
        else:
            outer = lambda: neuron.LIFNode(tau=tau, v_threshold=v_threshold, v_reset=v_reset,
                                           surrogate_function=surrogate.ATan(), detach_reset=True)
        self._build(self._cin0, outer, dict(v_threshold=v_threshold, v_reset=v_reset, use_plif=True, tau=tau),
                    multiply_factor,
                    neuron.IFNode(v_threshold=float('inf'), v_reset=v_reset, surrogate_function=surrogate.ATan()))


class fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike(
        fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike):
    """Monocular variant: 2 input channels, returns the depth list only (SNN_models.py:438-566)."""

    _cin0 = 2
    _returns_spikes = False

    def __init__(self, use_plif=False, detach_reset=True, tau=10., v_threshold=1.0, v_reset=0.0,
                 final_activation=nn.Identity, multiply_factor=1., in_channels=None):
        super().__init__(use_plif=use_plif, detach_reset=detach_reset, tau=tau, v_threshold=v_threshold,
                         v_reset=v_reset, multiply_factor=multiply_factor, in_channels=in_channels)
        self.final_activation = final_activation
