"""ORACLE -- test infrastructure only; never imported by the product path.

CPU restatement (plain torch, fp32) of the reference's analog comparison model:
    network/ANN_models.py:28-152  StereoSpike_equivalentANN
    network/blocks.py:41-83       ResBlock ('ADD')
    network/blocks.py:110-132     NNConvUpsampling
Conv -> activation -> BatchNorm everywhere, biased convolutions, four 1-channel heads accumulated by a non-firing IF pool
(SpikingJelly IFNode with v_threshold = inf: v <- v + x, never fires).  Same sub-module nesting, hence the same state-dict keys.
Pinned against the reference's own file executed by path in tests/test_oracle_vs_reference.py (build container only).
"""
import torch
import torch.nn as nn

from .ref_model import UpConv


class AnnResBlock(nn.Module):
    def __init__(self, c, act, bias=True):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=bias), act, nn.BatchNorm2d(c))
        self.conv2 = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=bias), act, nn.BatchNorm2d(c))

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        out += x                                    # blocks.py:73-74
        return out


class AnalogUNet(nn.Module):
    def __init__(self, activation_function=None):
        super().__init__()
        act = activation_function if activation_function is not None else nn.Sigmoid()
        enc = lambda ci, co, st: nn.Sequential(nn.Conv2d(ci, co, 5, st, 2, bias=True), act, nn.BatchNorm2d(co))
        dec = lambda ci, co, up: nn.Sequential(UpConv(ci, co, 5, up), act, nn.BatchNorm2d(co))
        self.bottom = enc(4, 32, 1)
        self.conv1, self.conv2, self.conv3, self.conv4 = enc(32, 64, 2), enc(64, 128, 2), enc(128, 256, 2), enc(256, 512, 2)
        self.bottleneck = nn.Sequential(AnnResBlock(512, act), AnnResBlock(512, act))
        self.deconv4, self.deconv3 = dec(512, 256, (33, 44)), dec(256, 128, (65, 87))
        self.deconv2, self.deconv1 = dec(128, 64, (130, 173)), dec(64, 32, (260, 346))
        self.predict_depth4 = nn.Sequential(UpConv(256, 1, 3, (260, 346), bias=True))
        self.predict_depth3 = nn.Sequential(UpConv(128, 1, 3, (260, 346), bias=True))
        self.predict_depth2 = nn.Sequential(UpConv(64, 1, 3, (260, 346), bias=True))
        self.predict_depth1 = nn.Sequential(UpConv(32, 1, 3, (260, 346), bias=True))
        self.v = 0.0                                # the I-neuron pool's potential (ANN_models.py:99)

    def reset(self):
        self.v = 0.0

    def forward(self, x):
        frame = x[:, 0]                             # ANN_models.py:103
        b = self.bottom(frame)
        c1 = self.conv1(b)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        c4 = self.conv4(c3)
        r = self.bottleneck(c4)
        a4 = self.deconv4(r) + c3
        self.v = self.v + self.predict_depth4(a4)
        d4 = self.v
        a3 = self.deconv3(a4) + c2
        self.v = self.v + self.predict_depth3(a3)
        d3 = self.v
        a2 = self.deconv2(a3) + c1
        self.v = self.v + self.predict_depth2(a2)
        d2 = self.v
        a1 = self.deconv1(a2) + b
        self.v = self.v + self.predict_depth1(a1)
        d1 = self.v
        return [d1, d2, d3, d4]


def randomize_batchnorm(model, seed=0):
    """Non-trivial BatchNorm affine parameters and running statistics (the defaults 1 / 0 / 0 / 1 would hide mix-ups)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.rand(m.bias.shape, generator=g) - 0.5)
                m.running_mean.copy_(torch.rand(m.running_mean.shape, generator=g) * 0.5 + 0.25)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.1 + 0.02)
