"""ORACLE -- test infrastructure only; never imported by the product path.

CPU restatement (plain PyTorch, fp32) of the slice of the third-party package
``spikingjelly.clock_driven`` that the reference's hot path calls.  SpikingJelly is
an un-vendored, un-pinned dependency of the reference (``/root/reference/requirements.txt:3``)
and is not installable here (no network), so its published single-step algorithm is
restated from the equations in SURVEY.md section 8(a) Row 6.

PARITY UNPINNED: the reference ships no tests, golden vectors or checkpoints for this
boundary, and SpikingJelly itself cannot be run here.  What pins this file instead:
  * the hand-derived known-answer vectors of SURVEY.md section 8(c) (tests/test_oracle_neurons.py);
  * the reference's own wiring executed *verbatim* on top of this shim
    (oracle/run_reference.py imports /root/reference/network/SNN_models.py by path).

Call sites in the reference that this file serves:
  network/blocks.py:8,150,157,175   (neuron.IFNode / ParametricLIFNode, surrogate.Sigmoid/ATan)
  network/SNN_models.py:6,24,26,78,150,266,338   (neuron.*, layer.Dropout, surrogate.*)
  train.py:12-13,118,221            (surrogate.ATan(), functional.reset_net)

Version choices (SpikingJelly generation 0.0.0.0.8 - 0.0.0.0.12, the one that still has
``clock_driven``):  heaviside is ``x >= 0``; ``surrogate.Sigmoid`` default alpha = 4.0;
``surrogate.ATan`` default alpha = 2.0; ``LIFNode`` charges with the decay-input form
``h = v + (x - (v - v_reset)) / tau``; hard reset; ``detach_reset`` detaches the spike
used in the reset only.
"""
import math
import sys
import types

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- surrogate
class _SigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        if x.requires_grad:
            ctx.save_for_backward(x)
            ctx.alpha = alpha
        return (x >= 0).to(x)

    @staticmethod
    def backward(ctx, grad_output):
        grad_x = None
        if ctx.needs_input_grad[0]:
            sgax = (ctx.saved_tensors[0] * ctx.alpha).sigmoid_()
            grad_x = grad_output * (1.0 - sgax) * sgax * ctx.alpha
        return grad_x, None


class _ATanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        if x.requires_grad:
            ctx.save_for_backward(x)
            ctx.alpha = alpha
        return (x >= 0).to(x)

    @staticmethod
    def backward(ctx, grad_output):
        grad_x = None
        if ctx.needs_input_grad[0]:
            x = ctx.saved_tensors[0]
            grad_x = ctx.alpha / 2 / (1 + (math.pi / 2 * ctx.alpha * x).pow_(2)) * grad_output
        return grad_x, None


class _SurrogateBase(nn.Module):
    def __init__(self, alpha, spiking=True):
        super().__init__()
        self.alpha = alpha
        self.spiking = spiking

    def forward(self, x):
        if self.spiking:
            return self._fn.apply(x, self.alpha)
        return self.primitive_function(x, self.alpha)


class Sigmoid(_SurrogateBase):
    _fn = _SigmoidFn

    def __init__(self, alpha=4.0, spiking=True):
        super().__init__(alpha, spiking)

    @staticmethod
    def primitive_function(x, alpha):
        return (x * alpha).sigmoid()


class ATan(_SurrogateBase):
    _fn = _ATanFn

    def __init__(self, alpha=2.0, spiking=True):
        super().__init__(alpha, spiking)

    @staticmethod
    def primitive_function(x, alpha):
        return (math.pi / 2 * alpha * x).atan_() / math.pi + 0.5


# ----------------------------------------------------------------------------- neurons
class BaseNode(nn.Module):
    """Single-step spiking neuron: charge -> fire -> (hard) reset.  State ``v`` is a python
    float (``v_reset``) until the first call, then a tensor; it persists between calls until
    ``reset()`` -- exactly the 'memory' behaviour of SpikingJelly's MemoryModule."""

    def __init__(self, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        super().__init__()
        self.v_threshold = v_threshold
        self.v_reset = v_reset
        self.detach_reset = detach_reset
        self.surrogate_function = surrogate_function if surrogate_function is not None else Sigmoid()
        self.v = 0.0 if v_reset is None else v_reset
        self.spike = 0.0

    def reset(self):
        self.v = 0.0 if self.v_reset is None else self.v_reset
        self.spike = 0.0

    def neuronal_charge(self, x):
        raise NotImplementedError

    def neuronal_fire(self):
        self.spike = self.surrogate_function(self.v - self.v_threshold)

    def neuronal_reset(self):
        spike = self.spike.detach() if self.detach_reset else self.spike
        if self.v_reset is None:
            self.v = self.v - spike * self.v_threshold
        else:
            self.v = (1.0 - spike) * self.v + spike * self.v_reset

    def forward(self, x):
        self.neuronal_charge(x)
        self.neuronal_fire()
        self.neuronal_reset()
        return self.spike


class IFNode(BaseNode):
    def neuronal_charge(self, x):
        self.v = self.v + x


class LIFNode(BaseNode):
    def __init__(self, tau=2.0, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        assert isinstance(tau, float) and tau > 1.0
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset)
        self.tau = tau

    def neuronal_charge(self, x):
        if self.v_reset is None or (isinstance(self.v_reset, float) and self.v_reset == 0.0):
            self.v = self.v + (x - self.v) / self.tau
        else:
            self.v = self.v + (x - (self.v - self.v_reset)) / self.tau


class ParametricLIFNode(BaseNode):
    def __init__(self, init_tau=2.0, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False):
        assert isinstance(init_tau, float) and init_tau > 1.0
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset)
        self.w = nn.Parameter(torch.as_tensor(-math.log(init_tau - 1.0)))

    def neuronal_charge(self, x):
        if self.v_reset is None or (isinstance(self.v_reset, float) and self.v_reset == 0.0):
            self.v = self.v + (x - self.v) * self.w.sigmoid()
        else:
            self.v = self.v + (x - (self.v - self.v_reset)) * self.w.sigmoid()


class Dropout(nn.Module):
    """Only referenced in an isinstance() check (SNN_models.py:26); never instantiated."""


def reset_net(net):
    for m in net.modules():
        if hasattr(m, 'reset'):
            m.reset()


# ----------------------------------------------------------------------------- module shim
def install_shim():
    """Register this file as ``spikingjelly.clock_driven.{neuron,surrogate,functional,layer,rnn}``
    so that the reference's own files import unmodified.  Idempotent."""
    if 'spikingjelly.clock_driven' in sys.modules and getattr(sys.modules['spikingjelly'], '_is_oracle_shim', False):
        return sys.modules['spikingjelly.clock_driven']
    sj = types.ModuleType('spikingjelly')
    sj._is_oracle_shim = True
    cd = types.ModuleType('spikingjelly.clock_driven')
    neuron = types.ModuleType('spikingjelly.clock_driven.neuron')
    for k in ('BaseNode', 'IFNode', 'LIFNode', 'ParametricLIFNode'):
        setattr(neuron, k, globals()[k])
    surrogate = types.ModuleType('spikingjelly.clock_driven.surrogate')
    surrogate.Sigmoid, surrogate.ATan = Sigmoid, ATan
    functional = types.ModuleType('spikingjelly.clock_driven.functional')
    functional.reset_net = reset_net
    layer = types.ModuleType('spikingjelly.clock_driven.layer')
    layer.Dropout = Dropout
    rnn = types.ModuleType('spikingjelly.clock_driven.rnn')
    for name, mod in (('neuron', neuron), ('surrogate', surrogate), ('functional', functional),
                      ('layer', layer), ('rnn', rnn)):
        setattr(cd, name, mod)
        sys.modules['spikingjelly.clock_driven.' + name] = mod
    sj.clock_driven = cd
    sys.modules['spikingjelly'] = sj
    sys.modules['spikingjelly.clock_driven'] = cd
    return cd
