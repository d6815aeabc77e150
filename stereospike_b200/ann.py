"""The analog comparison network of the reference (``network/ANN_models.py`` + the ANN blocks of ``network/blocks.py:15-83``):
the same encoder-decoder with biased convolutions, a classical activation and BatchNorm (Conv -> act -> BN), read out by
the same non-firing IF pool.  Same class names, constructor signatures, sub-module nesting and state-dict keys.

This is the paper's Table-4 baseline, not the spiking hot path: its activations are real numbers, so the int8 tensor-core
blocks do not apply.  Every convolution (plain strided Conv2d and nearest-upsample + conv) runs on the library's fp32
convolution kernels through the C ABI, forward and backward (``ss_conv_neuron_fwd`` as a non-firing step, ``ss_conv_dgrad``,
``ss_conv_wgrad``); the I-neuron pool on ``ss_neuron_fwd`` / ``ss_neuron_bwd``; the element-wise activation and BatchNorm are
left to torch.  Inputs must be CUDA tensors.
"""
import torch
import torch.nn as nn

from . import neuron, surrogate
from .blocks import NNConvUpsampling


def _run(seq, x):
    """nn.Sequential.forward with the Conv2d members routed to the library's convolution kernel."""
    from .engine import run_dense_conv
    for m in seq:
        x = run_dense_conv(m, x) if isinstance(m, nn.Conv2d) else m(x)
    return x


class BilinConvUpsampling(nn.Module):
    """Bilinear upsampling to ``up_size + (k-1)`` followed by a valid ``k x k`` convolution (blocks.py:15-38).  Not used by any
    model upstream (bilinear interpolation produces non-integer "spike counts"); kept for interface parity."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, up_size: tuple, bias: bool = False):
        super().__init__()
        self.up = nn.Sequential(
            nn.UpsamplingBilinear2d(size=(up_size[0] + (kernel_size - 1), up_size[1] + (kernel_size - 1))),
            nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=1, padding=0,
                      bias=bias),
        )

    def forward(self, x):
        return _run(self.up, x)


class ResBlock(nn.Module):
    """Standard residual block for ANNs (blocks.py:41-83): two Conv -> act -> BN stages, then the connect function."""

    def __init__(self, in_channels: int, connect_function='ADD', kernel_size: int = 3, bias: bool = False,
                 activation_function: nn.Module = None):
        super().__init__()
        if activation_function is None:
            activation_function = nn.Tanh()
        if connect_function not in ('ADD', 'MUL', 'AND', 'NMUL'):
            # upstream 'OR' raises too, and anything else hits a non-existent attribute (blocks.py:75-81)
            raise NotImplementedError(connect_function)
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size=kernel_size, stride=1, padding=(kernel_size - 1) // 2, bias=bias),
            activation_function,
            nn.BatchNorm2d(in_channels),
        )
        self.conv2 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size=kernel_size, stride=1, padding=(kernel_size - 1) // 2, bias=bias),
            activation_function,
            nn.BatchNorm2d(in_channels),
        )
        self.connect_function = connect_function

    def forward(self, x):
        identity = x
        out = _run(self.conv2, _run(self.conv1, x))
        if self.connect_function == 'ADD':
            out = out + identity
        elif self.connect_function in ('MUL', 'AND'):
            out = out * identity
        else:                                   # 'NMUL'
            out = identity * (1. - out)
        return out


class AnalogNet(nn.Module):
    """Book-keeping base class (ANN_models.py:9-25)."""

    def __init__(self):
        super().__init__()
        self.max_test_accuracy = float('inf')
        self.epoch = 0

    def increment_epoch(self):
        self.epoch += 1

    def get_max_accuracy(self):
        return self.max_test_accuracy

    def update_max_accuracy(self, new_acc):
        self.max_test_accuracy = new_acc

    def count_trainable_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class StereoSpike_equivalentANN(AnalogNet):
    """An analog network with exactly the architecture of StereoSpike: biases, BatchNorm and a classical activation instead of
    spiking neurons (ANN_models.py:28-152).  ``forward(x)``: x [B, frames, 4, 260, 346], frame 0 is used; returns the four
    depth maps [depth1 .. depth4] = the potential of the I-neuron pool after each head (stateful until reset, like upstream)."""

    def __init__(self, activation_function=None):
        super().__init__()
        act = activation_function if activation_function is not None else nn.Sigmoid()

        def enc(cin, cout, stride):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=5, stride=stride, padding=2, bias=True), act,
                                 nn.BatchNorm2d(cout))

        def dec(cin, cout, up_size):
            return nn.Sequential(NNConvUpsampling(cin, cout, 5, up_size), act, nn.BatchNorm2d(cout))

        self.bottom = enc(4, 32, 1)
        self.conv1 = enc(32, 64, 2)
        self.conv2 = enc(64, 128, 2)
        self.conv3 = enc(128, 256, 2)
        self.conv4 = enc(256, 512, 2)
        self.bottleneck = nn.Sequential(
            ResBlock(512, connect_function='ADD', bias=True, activation_function=act),
            ResBlock(512, connect_function='ADD', bias=True, activation_function=act),
        )
        self.deconv4 = dec(512, 256, (33, 44))
        self.deconv3 = dec(256, 128, (65, 87))
        self.deconv2 = dec(128, 64, (130, 173))
        self.deconv1 = dec(64, 32, (260, 346))
        self.predict_depth4 = nn.Sequential(NNConvUpsampling(256, 1, 3, (260, 346), bias=True))
        self.predict_depth3 = nn.Sequential(NNConvUpsampling(128, 1, 3, (260, 346), bias=True))
        self.predict_depth2 = nn.Sequential(NNConvUpsampling(64, 1, 3, (260, 346), bias=True))
        self.predict_depth1 = nn.Sequential(NNConvUpsampling(32, 1, 3, (260, 346), bias=True))
        self.Ineurons = neuron.IFNode(v_threshold=float('inf'), v_reset=0., surrogate_function=surrogate.ATan())

    def forward(self, x):
        frame = x[:, 0, :, :, :]
        out_bottom = _run(self.bottom, frame)
        out_conv1 = _run(self.conv1, out_bottom)
        out_conv2 = _run(self.conv2, out_conv1)
        out_conv3 = _run(self.conv3, out_conv2)
        out_conv4 = _run(self.conv4, out_conv3)
        out_rconv = self.bottleneck(out_conv4)

        out_add4 = _run(self.deconv4, out_rconv) + out_conv3
        self.Ineurons(self.predict_depth4(out_add4))
        depth4 = self.Ineurons.v
        out_add3 = _run(self.deconv3, out_add4) + out_conv2
        self.Ineurons(self.predict_depth3(out_add3))
        depth3 = self.Ineurons.v
        out_add2 = _run(self.deconv2, out_add3) + out_conv1
        self.Ineurons(self.predict_depth2(out_add2))
        depth2 = self.Ineurons.v
        out_add1 = _run(self.deconv1, out_add2) + out_bottom
        self.Ineurons(self.predict_depth1(out_add1))
        depth1 = self.Ineurons.v
        return [depth1, depth2, depth3, depth4]

    def set_init_depths_potentials(self, depth_prior):
        self.Ineurons.v = depth_prior


# the name upstream's package __init__ tries to import (a typo that makes `import network` fail at HEAD)
SteroSpike_equivalentANN = StereoSpike_equivalentANN
