"""Surrogate-gradient descriptors replacing ``spikingjelly.clock_driven.surrogate`` (reference call sites
train.py:118, network/blocks.py:142, network/SNN_models.py:12,266).  Forward is always the Heaviside step
(spike when ``h - v_th >= 0``); the object only selects the derivative used by the fused backward kernel
(``ss_neuron_bwd``): ATan: alpha/2 / (1 + (pi/2*alpha*u)^2), Sigmoid: alpha*s(alpha*u)*(1 - s(alpha*u)).
"""
import torch.nn as nn

from ._lib import SS_SURR_ATAN, SS_SURR_SIGMOID


class SurrogateFunctionBase(nn.Module):
    kind = None

    def __init__(self, alpha, spiking=True):
        super().__init__()
        if not spiking:
            raise NotImplementedError('stereospike_b200 surrogates are spiking-only (Heaviside forward)')
        self.alpha = float(alpha)
        self.spiking = True

    def extra_repr(self):
        return f'alpha={self.alpha}, spiking={self.spiking}'


class ATan(SurrogateFunctionBase):
    kind = SS_SURR_ATAN

    def __init__(self, alpha=2.0, spiking=True):
        super().__init__(alpha, spiking)


class Sigmoid(SurrogateFunctionBase):
    kind = SS_SURR_SIGMOID

    def __init__(self, alpha=4.0, spiking=True):
        super().__init__(alpha, spiking)
