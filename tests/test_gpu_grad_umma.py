"""Tensor-core gradient kernels (ss_corr_bf16, ss_conv_wgrad_bf16) against PyTorch autograd of the same op, evaluated in
float64 so that the reference itself carries no algorithm-dependent error (cuDNN's fp32 FFT / Winograd gradients are only
good to ~1e-4).  Operands are rounded to bf16 first, so the two sides differ only by the kernels' fp32 accumulation
(tolerance 2e-4 relative to the largest gradient); a second check against the un-rounded gradient bounds the bf16 error
(cosine >= 0.9999)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # kind, Cin, Cout, ks, Hin, Win, stride, pad, up, T, B
    ('conv', 64, 64, 3, 17, 22, 1, 1, None, 2, 3),       # bottleneck geometry
    ('conv', 32, 64, 5, 37, 45, 2, 2, None, 2, 2),       # strided encoder conv (odd sizes)
    ('conv', 64, 32, 5, 40, 30, 2, 2, None, 3, 1),       # strided, even sizes, 32 output channels (4 shifted copies in wgrad)
    ('conv', 32, 32, 5, 21, 19, 2, 2, None, 1, 2),       # destination 32 channels (ntile 32)
    ('upconv', 64, 32, 5, 17, 22, 1, 0, (33, 44), 2, 2),  # decoder geometry
    ('upconv', 128, 64, 5, 9, 12, 1, 0, (20, 23), 5, 1),
    ('conv', 512, 512, 3, 17, 22, 1, 1, None, 5, 2),     # deep: streamed weights, several channel blocks
    ('upconv', 64, 16, 5, 11, 13, 1, 0, (23, 27), 2, 1),  # 16 output channels (4 shifted copies, half-empty rows)
    ('conv', 32, 48, 5, 19, 26, 1, 2, None, 2, 2),       # 48 output channels, same-padded 5x5 (2 shifted copies)
    ('conv', 256, 160, 5, 20, 18, 2, 2, None, 2, 1),     # 160 output channels: second 128-block half empty
]


def _case(kind, Cin, Cout, ks, Hin, Win, stride, pad, up, T, B, seed=0):
    from stereospike_b200 import ops
    g = torch.Generator().manual_seed(seed)
    if kind == 'conv':
        Hout, Wout = ops.conv_out_size(Hin, ks, stride, pad), ops.conv_out_size(Win, ks, stride, pad)
        geom = ops.BlockGeom('conv', Cin, Cout, ks, Hin, Win, Hout, Wout, stride, pad)
    else:
        Hout, Wout = up
        geom = ops.BlockGeom('upconv', Cin, Cout, ks, Hin, Win, Hout, Wout)
    x = ((torch.rand(T, B, Cin, Hin, Win, generator=g) < 0.2).float() * torch.randint(1, 4, (T, B, Cin, Hin, Win), generator=g)).float()
    w = (torch.rand(Cout, Cin, ks, ks, generator=g) * 2 - 1) / (Cin * ks * ks) ** 0.5
    gy = torch.randn(T, B, Cout, Hout, Wout, generator=g) * (torch.rand(T, B, Cout, Hout, Wout, generator=g) < 0.5)
    return geom, x, w, gy


def _autograd(geom, x, w, gy):
    T, B = x.shape[:2]
    xi = x.double().reshape(T * B, *x.shape[2:]).clone().requires_grad_(True)
    wi = w.double().clone().requires_grad_(True)
    if geom.kind == 'upconv':
        y = F.conv2d(F.interpolate(xi, size=(geom.Hout + geom.ks - 1, geom.Wout + geom.ks - 1), mode='nearest'), wi)
    else:
        y = F.conv2d(xi, wi, stride=geom.stride, padding=geom.pad)
    y.backward(gy.double().reshape(T * B, *gy.shape[2:]))
    return xi.grad.reshape(x.shape).float(), wi.grad.float()


def _cos(a, b):
    return float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm() + 1e-300))


@pytest.mark.parametrize('case', CASES, ids=lambda c: f'{c[0]}-{c[1]}to{c[2]}-k{c[3]}s{c[6]}')
def test_dgrad_tensor_core(case):
    from stereospike_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    geom, x, w, gy = _case(*case)
    T, B = x.shape[:2]
    dev = torch.device('cuda')
    gy16 = gy.bfloat16()
    w16 = w.bfloat16()
    gx_ref, _ = _autograd(geom, x.cuda(), w16.float().cuda(), gy16.float().cuda())       # same rounded operands
    gx_full, _ = _autograd(geom, x.cuda(), w.cuda(), gy.cuda())
    plan = ops.DgradPlan(w.to(dev), geom, dev)
    g_dev = gy16.permute(0, 1, 3, 4, 2).contiguous().to(dev)
    seed_val = 0.25
    dst = torch.full((T, B, geom.Hin, geom.Win, geom.Cin), seed_val, dtype=torch.float32, device=dev)   # accumulate semantics
    plan.run(g_dev, dst, T, B)
    torch.cuda.synchronize()
    got = (dst - seed_val).permute(0, 1, 4, 2, 3)
    scale = float(gx_ref.abs().max())
    assert float((got - gx_ref).abs().max()) <= 2e-4 * scale + 1e-6, (float((got - gx_ref).abs().max()), scale)
    assert _cos(got, gx_full) >= 0.9999


@pytest.mark.parametrize('case', CASES, ids=lambda c: f'{c[0]}-{c[1]}to{c[2]}-k{c[3]}s{c[6]}')
def test_wgrad_tensor_core(case):
    from stereospike_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    geom, x, w, gy = _case(*case)
    T, B = x.shape[:2]
    dev = torch.device('cuda')
    gy16 = gy.bfloat16()
    _, gw_ref = _autograd(geom, x.cuda(), w.cuda(), gy16.float().cuda())
    _, gw_full = _autograd(geom, x.cuda(), w.cuda(), gy.cuda())
    xb = x.permute(0, 1, 3, 4, 2).contiguous().to(dev, torch.uint8)
    g_dev = gy16.permute(0, 1, 3, 4, 2).contiguous().to(dev)
    g_wkn = ops.conv_wgrad_bf16(xb, g_dev, geom, T, B)
    torch.cuda.synchronize()
    got = ops.kn_to_weight(g_wkn, geom.Cout, geom.Cin, geom.ks)
    scale = float(gw_ref.abs().max())
    assert float((got - gw_ref).abs().max()) <= 2e-4 * scale + 1e-6, (float((got - gw_ref).abs().max()), scale)
    assert _cos(got, gw_full) >= 0.9999


@pytest.mark.parametrize('cin', [32, 64])
def test_wgrad_full_range_inputs(cin):
    """Event-count frames as a multi-channel first layer (channel-concatenated temporal mode, reference train.py:206-218): x holds
    any u8 value, which needs the full-range u8 -> bf16 conversion of the patch producers (x_full_range=True); the default
    conversion is only defined below 128."""
    from stereospike_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    geom, x, w, gy = _case('conv', cin, 64, 5, 20, 30, 1, 2, None, 2, 2)
    T, B = x.shape[:2]
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, 256, x.shape, generator=g).float() * (torch.rand(x.shape, generator=g) < 0.3)
    assert float(x.max()) > 200
    dev = torch.device('cuda')
    gy16 = gy.bfloat16()
    _, gw_ref = _autograd(geom, x.cuda(), w.cuda(), gy16.float().cuda())
    xb = x.permute(0, 1, 3, 4, 2).contiguous().to(dev, torch.uint8)
    g_dev = gy16.permute(0, 1, 3, 4, 2).contiguous().to(dev)
    got = ops.kn_to_weight(ops.conv_wgrad_bf16(xb, g_dev, geom, T, B, x_full_range=True), geom.Cout, geom.Cin, geom.ks)
    torch.cuda.synchronize()
    scale = float(gw_ref.abs().max())
    assert float((got - gw_ref).abs().max()) <= 2e-4 * scale + 1e-6, (float((got - gw_ref).abs().max()), scale)
