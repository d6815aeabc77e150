"""ORACLE -- test infrastructure only; never imported by the product path.

CPU (or eager-GPU) restatement in plain PyTorch fp32 of the reference's spiking U-Net:
  network/blocks.py:90-107    MultiplyBy
  network/blocks.py:110-132   NNConvUpsampling  (UpsamplingNearest2d(size=up+k-1) -> Conv2d(k, s1, p0))
  network/blocks.py:135-181   SEWResBlock ('ADD' connect function only)
  network/SNN_models.py:63-192    StereoSpike              (IF neurons, binocular)
  network/SNN_models.py:251-376   fromZero_..._Matt_SpikeFlowNetLike   (LIF / PLIF, binocular)
  network/SNN_models.py:438-566   fromZero_..._monocular_SpikeFlowNetLike (LIF / PLIF, Cin=2,
                                  returns the depth list only)
Module nesting reproduces the reference's state-dict keys.

PARITY UNPINNED by the reference (it has no tests / vectors / weights).  This restatement is
pinned instead by oracle/run_reference.py, which executes the reference's own files by
path on top of oracle/sj_compat.py and must agree with this file bit-for-bit on CPU
(tests/test_oracle_vs_reference.py), and by tests/golden/*.npz generated from that run.

``forward_seq`` implements the T-loop contract of SURVEY.md section 0: the single-frame module
called once per timestep with no reset in between.
"""
import torch
import torch.nn as nn

from . import sj_compat as sj

H0, W0 = 260, 346
# (name, cin, cout) of the strided encoder convs; sizes follow from 5x5 / s2 / p2 on 260x346
ENC = (('conv1', 32, 64), ('conv2', 64, 128), ('conv3', 128, 256), ('conv4', 256, 512))
# (name, cin, cout, up_size) of the decoder
DEC = (('deconv4', 512, 256, (33, 44)), ('deconv3', 256, 128, (65, 87)),
       ('deconv2', 128, 64, (130, 173)), ('deconv1', 64, 32, (260, 346)))
HEADS = (('predict_depth4', 256), ('predict_depth3', 128), ('predict_depth2', 64), ('predict_depth1', 32))


class Gain(nn.Module):
    """blocks.py:90-107 with learnable=False (no call site sets it)."""

    def __init__(self, g):
        super().__init__()
        self.scale_value = g

    def forward(self, x):
        return torch.mul(x, self.scale_value)


class UpConv(nn.Module):
    """blocks.py:110-132; attribute ``up`` keeps the key ``<name>.0.up.1.weight``."""

    def __init__(self, cin, cout, k, up_size, bias=False):
        super().__init__()
        self.up = nn.Sequential(
            nn.UpsamplingNearest2d(size=(up_size[0] + k - 1, up_size[1] + k - 1)),
            nn.Conv2d(cin, cout, k, stride=1, padding=0, bias=bias))

    def forward(self, x):
        return self.up(x)


class SEWBlock(nn.Module):
    """blocks.py:135-171 (ADD)."""

    def __init__(self, c, make_neuron, gain):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=False), Gain(gain))
        self.sn1 = make_neuron()
        self.conv2 = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=False), Gain(gain))
        self.sn2 = make_neuron()

    def forward(self, x):
        out = self.sn2(self.conv2(self.sn1(self.conv1(x))))
        out += x
        return out


class SpikingUNet(nn.Module):
    """variant: 'if' (StereoSpike), 'lif', 'plif' (the two fromZero classes; ``monocular`` picks Cin=2)."""

    def __init__(self, variant='if', monocular=False, surrogate_function=None, tau=10.0,
                 v_threshold=1.0, v_reset=0.0, multiply_factor=1.0, in_channels=None):
        super().__init__()
        assert variant in ('if', 'lif', 'plif')
        self.variant, self.monocular = variant, monocular
        g = multiply_factor
        if variant == 'if':
            # SNN_models.py:71-72: v_threshold / v_reset are NOT forwarded -> always 1.0 / 0.0
            v_threshold, v_reset = 1.0, 0.0
            sf = surrogate_function if surrogate_function is not None else sj.Sigmoid()
            self.surrogate_fct = sf
            outer = lambda: sj.IFNode(v_threshold, v_reset, sf, detach_reset=True)
            # SNN_models.py:105-106 pass no surrogate -> SEWResBlock default Sigmoid (blocks.py:142)
            inner = lambda: sj.IFNode(v_threshold, v_reset, sj.Sigmoid(), detach_reset=True)
            i_sf = sf
        elif variant == 'lif':
            outer = lambda: sj.LIFNode(tau, v_threshold, v_reset, sj.ATan(), detach_reset=True)
            # SNN_models.py:293-294: bottleneck hard-codes use_plif=True
            inner = lambda: sj.ParametricLIFNode(tau, v_threshold, v_reset, sj.Sigmoid(), detach_reset=True)
            i_sf = sj.ATan()
        else:
            outer = lambda: sj.ParametricLIFNode(tau, v_threshold, v_reset, None, detach_reset=True)
            inner = lambda: sj.ParametricLIFNode(tau, v_threshold, v_reset, sj.Sigmoid(), detach_reset=True)
            i_sf = sj.ATan()

        # in_channels: the hand edit train.py:206-213 asks for in the channel-concatenated temporal mode ("number of filters in the
        # first convolution should be changed accordingly"): 2 * nfpdm * cameras input channels
        cin0 = in_channels if in_channels is not None else (2 if monocular else 4)
        self.bottom = nn.Sequential(nn.Conv2d(cin0, 32, 5, 1, 2, bias=False), Gain(g), outer())
        for name, ci, co in ENC:
            setattr(self, name, nn.Sequential(nn.Conv2d(ci, co, 5, 2, 2, bias=False), Gain(g), outer()))
        self.bottleneck = nn.Sequential(SEWBlock(512, inner, g), SEWBlock(512, inner, g))
        for name, ci, co, up in DEC:
            setattr(self, name, nn.Sequential(UpConv(ci, co, 5, up), Gain(g), outer()))
        for name, ci in HEADS:
            setattr(self, name, nn.Sequential(UpConv(ci, 1, 3, (H0, W0), bias=True), Gain(g)))
        self.Ineurons = sj.IFNode(float('inf'), 0.0 if variant == 'if' else v_reset, i_sf)

    def forward(self, x, return_all=False):
        frame = x[:, 0]
        b = self.bottom(frame)
        c1 = self.conv1(b)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        c4 = self.conv4(c3)
        r = self.bottleneck(c4)
        depths, adds, decs = [], [], []
        cur = r
        for (name, _, _, _), skip, (hname, _) in zip(DEC, (c3, c2, c1, b), HEADS):
            d = getattr(self, name)(cur)
            cur = d + skip
            self.Ineurons(getattr(self, hname)(cur))
            depths.append(self.Ineurons.v)
            adds.append(cur)
            decs.append(d)
        depths = depths[::-1]                      # [depth1, depth2, depth3, depth4]
        spks = [r] + adds                          # [out_rconv, out_add4, out_add3, out_add2, out_add1]
        if return_all:
            layers = dict(out_bottom=b, out_conv1=c1, out_conv2=c2, out_conv3=c3, out_conv4=c4, out_rconv=r,
                          out_deconv4=decs[0], out_add4=adds[0], out_deconv3=decs[1], out_add3=adds[1],
                          out_deconv2=decs[2], out_add2=adds[2], out_deconv1=decs[3], out_add1=adds[3])
            return depths, spks, layers
        if self.monocular:
            return depths
        return depths, spks

    def forward_seq(self, x_seq, return_all=False):
        """x_seq [B,T,C,260,346]; state must have been reset by the caller (train.py:221)."""
        out = None
        for t in range(x_seq.shape[1]):
            out = self.forward(x_seq[:, t:t + 1], return_all=return_all)
        return out


def mean_depth_error(pred, gt):
    """network/metrics.py:83-95 -- NaN-masked mean absolute error."""
    mask = ~torch.isnan(gt)
    return (pred - gt)[mask].abs().sum() / mask.count_nonzero()


def firing_rates(layers):
    """SNN_models.py:194-245: density = count_nonzero / numel per layer."""
    return {k: float(v.count_nonzero()) / v.numel() for k, v in layers.items()}


def synthetic_inputs(B, T, cin=4, lam=0.05, seed=0):
    """SURVEY.md section 8(d) recipe: Poisson event counts, ~5 % active pixels, bf16-exact."""
    g = torch.Generator().manual_seed(seed)
    return torch.poisson(torch.full((B, T, cin, H0, W0), lam), generator=g)


def synthetic_label(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    lab = torch.rand(B, 1, H0, W0, generator=g) * 10
    holes = torch.rand(B, 1, H0, W0, generator=g) < 0.2
    lab[holes] = float('nan')
    return lab
