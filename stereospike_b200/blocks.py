"""Drop-in replacements for the spiking blocks of the reference's ``network/blocks.py`` (same class names,
constructor signatures, sub-module nesting and therefore state-dict keys):

    MultiplyBy        blocks.py:90-107
    NNConvUpsampling  blocks.py:110-132   (key ``up.1.weight`` / ``up.1.bias``)
    SEWResBlock       blocks.py:135-181   ('ADD' connect function; keys ``conv{1,2}.0.weight``, ``sn{1,2}.w``)

Inside the models these modules hold the parameters and neuron state; the arithmetic runs in the fused CUDA
kernels (stereospike_b200/engine.py).  Called on their own they take / return the reference's NCHW fp32
tensors, one timestep per call, stateful, and run the same kernels through a one-block engine.
"""
import torch
import torch.nn as nn

from . import neuron, surrogate


class MultiplyBy(nn.Module):
    """By multiplying input values by a certain parameter, it should allow subsequent PLIFNodes to actually spike and
    solve the vanishing spike phenomenon (blocks.py:90-97)."""

    def __init__(self, scale_value: float = 5., learnable: bool = False):
        super().__init__()
        if learnable:
            self.scale_value = nn.Parameter(torch.Tensor([scale_value]))
        else:
            self.scale_value = scale_value

    def forward(self, x):
        # a scalar gain; inside the models it is folded into the conv epilogue of the fused kernels
        return torch.mul(x, self.scale_value)

    def gain(self):
        if isinstance(self.scale_value, nn.Parameter) and self.scale_value.requires_grad and torch.is_grad_enabled():
            # the fused kernels take the gain as a launch constant: a learnable gain would silently get no gradient
            raise NotImplementedError('MultiplyBy(learnable=True) cannot be trained inside the fused blocks (no call site of the '
                                      'reference sets it); freeze it (requires_grad_(False)) or run under torch.no_grad()')
        return float(self.scale_value)


class NNConvUpsampling(nn.Module):
    """Nearest-neighbour upsampling to ``up_size + (k-1)`` followed by a valid ``k x k`` convolution
    (blocks.py:110-132).  The fused kernels never materialise the upsampled tensor: the (ymap, xmap) gather
    tables of stereospike_b200.ops.upsample_axis_map index the low-resolution source directly."""

    def __init__(self, in_channels, out_channels, kernel_size, up_size, bias=False):
        super().__init__()
        self.up = nn.Sequential(
            nn.UpsamplingNearest2d(size=(up_size[0] + (kernel_size - 1), up_size[1] + (kernel_size - 1))),
            nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=1,
                      padding=0, bias=bias),
        )
        self.up_size = (int(up_size[0]), int(up_size[1]))

    def forward(self, x):
        from .engine import run_linear_block
        return run_linear_block(self, x)


class SEWResBlock(nn.Module):
    """Spike-Element-Wise residual block, 'ADD' only (blocks.py:135-171):
    out = sn2(g*conv2(sn1(g*conv1(x)))) + x."""

    def __init__(self, in_channels: int, connect_function='ADD', v_threshold=1., v_reset=0.,
                 surrogate_function=None, use_plif=False, tau=2., multiply_factor=1.):
        super().__init__()
        if connect_function != 'ADD':
            # every reference model uses 'ADD'; the other branches reference a non-existent attribute upstream
            raise NotImplementedError(connect_function)
        if surrogate_function is None:
            surrogate_function = surrogate.Sigmoid()
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size=3, padding=1, stride=1, bias=False),
            MultiplyBy(multiply_factor),
        )
        self.sn1 = neuron.ParametricLIFNode(init_tau=tau, v_threshold=v_threshold, v_reset=v_reset,
                                            surrogate_function=surrogate_function, detach_reset=True) if use_plif \
            else neuron.IFNode(v_threshold=v_threshold, v_reset=v_reset, surrogate_function=surrogate_function,
                               detach_reset=True)
        self.conv2 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size=3, padding=1, stride=1, bias=False),
            MultiplyBy(multiply_factor),
        )
        self.sn2 = neuron.ParametricLIFNode(init_tau=tau, v_threshold=v_threshold, v_reset=v_reset,
                                            surrogate_function=surrogate_function, detach_reset=True) if use_plif \
            else neuron.IFNode(v_threshold=v_threshold, v_reset=v_reset, surrogate_function=surrogate_function,
                               detach_reset=True)
        self.connect_function = connect_function

    def forward(self, x):
        from .engine import run_sew_block
        return run_sew_block(self, x)
