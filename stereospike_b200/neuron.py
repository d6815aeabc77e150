"""Spiking neuron modules replacing ``spikingjelly.clock_driven.neuron`` (IFNode / LIFNode / ParametricLIFNode;
reference call sites network/blocks.py:150,157 and network/SNN_models.py:78,150,266).

Inside the models these modules are *descriptors + state holders*: the fused block kernels read their
hyper-parameters and carry the membrane potential.  Called on their own (``node(x)``) they run the stand-alone
CUDA neuron kernel (``ss_neuron_fwd`` / ``ss_neuron_bwd``), single step, stateful -- the SpikingJelly contract:
``v`` is the python float ``v_reset`` until the first call, then a tensor shaped like the input, until ``reset()``.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib, surrogate
from ._lib import SS_NEURON_IF, SS_NEURON_LIF, SS_NEURON_PLIF
from .ops import _ptr, _require_cuda, _stream


class _NeuronStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v0, decay, node):
        xc = x.contiguous().float()
        N = xc.numel()
        v = v0.detach().contiguous().float().clone()
        s = torch.empty_like(xc)
        h = torch.empty_like(xc)
        rc = _lib.lib().ss_neuron_fwd(1, N, node.kind, node.v_threshold, node._v_reset_value(), node._tau_value(),
                                      _ptr(decay), _ptr(xc), _ptr(v), _ptr(s), _ptr(h), _stream())
        _lib.check(rc, 'ss_neuron_fwd')
        ctx.node = node
        ctx.save_for_backward(h, v0.detach().contiguous().float(), decay if decay is not None else torch.empty(0))
        return s, v

    @staticmethod
    def backward(ctx, g_s, g_v):
        # Both outputs are differentiable, like upstream: the potential of a non-firing pool IS the prediction (Ineurons.v,
        # network/SNN_models.py:150, ANN_models.py:99) and the state carried to the next call keeps its graph until detach().
        node = ctx.node
        h, v0, decay = ctx.saved_tensors
        decay = decay if decay.numel() else None
        g_s = g_s.contiguous().float()
        g_v = g_v.contiguous().float() if g_v is not None else None
        g_x = torch.empty_like(h)
        g_v0 = torch.empty_like(h) if ctx.needs_input_grad[1] else None
        g_decay = torch.zeros((1,), dtype=torch.float32, device=h.device) if decay is not None else None     # decay_tensor() is [1]
        sf = node.surrogate_function
        rc = _lib.lib().ss_neuron_bwd(1, h.numel(), node.kind, sf.kind, sf.alpha, 1.0, node.v_threshold,
                                      node._v_reset_value(), node._tau_value(), _ptr(decay), _ptr(h), _ptr(v0),
                                      _ptr(g_s), _ptr(g_v), _ptr(g_x), _ptr(g_v0), _ptr(g_decay), _stream())
        _lib.check(rc, 'ss_neuron_bwd')
        return g_x, g_v0, g_decay, None


class BaseNode(nn.Module):
    kind = None

    def __init__(self, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        super().__init__()
        if v_reset is None:
            raise NotImplementedError('soft reset (v_reset=None) is not used by the reference and not implemented')
        self.v_threshold = float(v_threshold)
        self.v_reset = float(v_reset)
        self.detach_reset = detach_reset
        self.surrogate_function = surrogate_function if surrogate_function is not None else surrogate.Sigmoid()
        self.v = self.v_reset
        self.spike = 0.0

    def _v_reset_value(self):
        return self.v_reset

    def _tau_value(self):
        return 2.0

    def decay_tensor(self):
        return None

    def reset(self):
        self.v = self.v_reset
        self.spike = 0.0

    def extra_repr(self):
        return f'v_threshold={self.v_threshold}, v_reset={self.v_reset}, detach_reset={self.detach_reset}'

    def forward(self, x):
        _require_cuda(x, 'x')
        v0 = self.v if isinstance(self.v, torch.Tensor) else torch.full_like(x, self.v, dtype=torch.float32)
        s, v = _NeuronStep.apply(x, v0, self.decay_tensor(), self)
        if not self.detach_reset and x.requires_grad and math.isfinite(self.v_threshold):
            # the fused backward implements the reference's detach_reset=True contract only
            raise NotImplementedError('detach_reset=False with a finite threshold is not implemented')
        self.v = v.view_as(x)
        self.spike = s.view_as(x)
        return self.spike


class IFNode(BaseNode):
    kind = SS_NEURON_IF


class LIFNode(BaseNode):
    kind = SS_NEURON_LIF

    def __init__(self, tau=2.0, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        assert isinstance(tau, float) and tau > 1.0
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset)
        self.tau = tau

    def _tau_value(self):
        return self.tau

    def extra_repr(self):
        return super().extra_repr() + f', tau={self.tau}'


class ParametricLIFNode(BaseNode):
    kind = SS_NEURON_PLIF

    def __init__(self, init_tau=2.0, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        assert isinstance(init_tau, float) and init_tau > 1.0
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset)
        self.w = nn.Parameter(torch.as_tensor(-math.log(init_tau - 1.0)))

    def decay_tensor(self):
        # 1/tau = sigmoid(w): a one-element device tensor; torch autograd carries d sigmoid / d w
        if torch.is_grad_enabled() and self.w.requires_grad:
            return self.w.sigmoid().reshape(1).float()
        # inference: one tiny launch per PLIF layer and call otherwise (4 of the ~30 launches of a forward); cached on the
        # parameter's version counter (optimizer steps and load_state_dict bump it)
        key = (self.w.data_ptr(), self.w._version, self.w.device)
        c = getattr(self, '_decay_cache', None)
        if c is None or c[0] != key:
            with torch.no_grad():
                c = (key, self.w.sigmoid().reshape(1).float())
            object.__setattr__(self, '_decay_cache', c)
        return c[1]

    def extra_repr(self):
        with torch.no_grad():
            return super().extra_repr() + f', tau={1.0 / float(self.w.sigmoid())}'
