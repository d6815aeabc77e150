"""Tensor-level wrappers over the C ABI: geometry tables, weight packing, fused block / heads calls.

PyTorch is used here for device memory, streams and autograd plumbing only; all arithmetic of the hot path
happens inside libstereospike_b200.so.
"""
import ctypes
import functools

import numpy as np
import torch

from . import _lib
from ._lib import (SS_IMPL_AUTO, SS_IMPL_SIMT, SS_IMPL_UMMA, SS_IN_U8_TBHWC, SS_IN_F32_BTCHW,
                   SS_NEURON_IF, SS_NEURON_LIF, SS_NEURON_PLIF)

ACT_DTYPE = torch.uint8          # spikes / spike sums / event counts in HBM: u8 NHWC, exact


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'stereospike_b200: `{name}` must be a CUDA tensor -- the hot path has no CPU fallback')


# ----------------------------------------------------------------------------------------- geometry tables
def conv_axis_map(n_in, n_out, ks, stride, pad):
    """Source index read by output o at tap k of a zero-padded strided conv (-1 = padding)."""
    o = np.arange(n_out)[:, None]
    k = np.arange(ks)[None, :]
    s = o * stride - pad + k
    s = np.where((s >= 0) & (s < n_in), s, -1)
    return s.astype(np.int32)


def upsample_axis_map(n_in, n_out, ks):
    """Source index read by output o at tap k of  UpsamplingNearest2d(size=n_out+ks-1) -> valid conv
    (reference network/blocks.py:124-127).  Index rule of ATen's nearest kernel: float32
    ``min(floor(dst * (float(in) / out)), in - 1)``."""
    n_up = n_out + ks - 1
    scale = np.float32(n_in) / np.float32(n_up)
    dst = np.arange(n_up, dtype=np.float32)
    src = np.minimum(np.floor(dst * scale).astype(np.int64), n_in - 1)
    o = np.arange(n_out)[:, None]
    k = np.arange(ks)[None, :]
    return src[o + k].astype(np.int32)


@functools.lru_cache(maxsize=None)
def _maps_cached(kind, Hin, Win, Hout, Wout, ks, stride, pad, device):
    if kind == 'conv':
        ym, xm = conv_axis_map(Hin, Hout, ks, stride, pad), conv_axis_map(Win, Wout, ks, stride, pad)
    else:
        ym, xm = upsample_axis_map(Hin, Hout, ks), upsample_axis_map(Win, Wout, ks)
    dev = torch.device(device)
    return (torch.from_numpy(ym.reshape(-1).copy()).to(dev), torch.from_numpy(xm.reshape(-1).copy()).to(dev))


def conv_out_size(n, ks, stride, pad):
    return (n + 2 * pad - ks) // stride + 1


class BlockGeom:
    """Static geometry of one fused block (everything but T, B and the pointers)."""

    def __init__(self, kind, Cin, Cout, ks, Hin, Win, Hout, Wout, stride=1, pad=0):
        assert kind in ('conv', 'upconv')
        self.kind, self.Cin, self.Cout, self.ks = kind, Cin, Cout, ks
        self.Hin, self.Win, self.Hout, self.Wout = Hin, Win, Hout, Wout
        self.stride, self.pad = stride, pad
        self.K = ks * ks * Cin

    def maps(self, device):
        return _maps_cached(self.kind, self.Hin, self.Win, self.Hout, self.Wout, self.ks, self.stride, self.pad,
                            str(device))


# ----------------------------------------------------------------------------------------- weight packing
def weight_to_kn(weight):
    """OIHW fp32 (the reference's state-dict layout) -> [K][Cout], k = (ky*ks + kx)*Cin + c."""
    co, ci, kh, kw = weight.shape
    return weight.detach().permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous().float()


def kn_to_weight(w_kn, co, ci, ks):
    return w_kn.reshape(ks, ks, ci, co).permute(3, 2, 0, 1).contiguous()


def pack_weights_i8(weight, planes, cin_pad=None):
    """OIHW fp32 -> (int8 digit planes in the kernel's shared-memory image, wscale fp32 [Cout], wexp int32 [Cout]).
    ``cin_pad``: zero-pad the input channels up to this count (the first layer's 2 channels -> 4)."""
    w = weight.detach().float()
    co, ci, kh, kw = w.shape
    if cin_pad is not None and cin_pad > ci:
        w = torch.cat([w, w.new_zeros(co, cin_pad - ci, kh, kw)], dim=1)
        ci = cin_pad
    w = w.contiguous()
    nbytes = co * 128 * planes if ci <= 4 else co * ci * kh * kw * planes
    out = torch.empty(nbytes, dtype=torch.int8, device=w.device)
    wscale = torch.empty(co, dtype=torch.float32, device=w.device)
    wexp = torch.empty(co, dtype=torch.int32, device=w.device)
    _lib.check(_lib.lib().ss_pack_weights_i8(_ptr(w), co, ci, kh, planes, _ptr(out), _ptr(wscale), _ptr(wexp), _stream()),
               'ss_pack_weights_i8')
    return out, wscale, wexp


def pack_events(x_seq, status=None):
    """fp32 [B,T,C,H,W] event-count frames -> u8 [T,B,H,W,4] (the first block's tensor-core input)."""
    _require_cuda(x_seq, 'x')
    B, T, C, H, W = x_seq.shape
    out = torch.empty((T, B, H, W, 4), dtype=torch.uint8, device=x_seq.device)
    _lib.check(_lib.lib().ss_pack_events(_ptr(x_seq), B, T, C, H, W, _ptr(out), _ptr(status), _stream()), 'ss_pack_events')
    return out


def _check_block_io(x, g, T, B, resid, v_in, decay, out_shape):
    if resid is not None:
        assert resid.dtype == ACT_DTYPE and resid.is_contiguous() and tuple(resid.shape) == out_shape
    if v_in is not None:
        assert v_in.dtype == torch.float32 and v_in.is_contiguous() and tuple(v_in.shape) == (B, g.Hout, g.Wout, g.Cout)
    if decay is not None:
        assert decay.dtype == torch.float32 and decay.numel() == 1 and decay.is_cuda


# ----------------------------------------------------------------------------------------- fused block
def conv_i8_fwd(x, geom, w_i8, wscale, *, T, B, neuron, gain, v_th, v_reset, tau=2.0, decay=None, v_in=None,
                want_v_out=False, resid=None, want_h=False, planes=3, cin=None, tsum=None):
    """Tensor-core fused block over all T timesteps (ss_conv_i8_fwd).  x: u8 [T,B,Hin,Win,Cin].
    ``tsum``: optional u8 [B,Hout,Wout,Cout] receiving the sum of the first T-1 output steps (input of the linear heads).
    Returns (out u8 [T,B,Hout,Wout,Cout], v_out, h_seq)."""
    _require_cuda(x, 'x')
    dev = x.device
    g = geom
    cin = g.Cin if cin is None else cin
    assert x.dtype == ACT_DTYPE and x.is_contiguous() and tuple(x.shape) == (T, B, g.Hin, g.Win, cin), \
        (x.dtype, tuple(x.shape), (T, B, g.Hin, g.Win, cin))
    out_shape = (T, B, g.Hout, g.Wout, g.Cout)
    out = torch.empty(out_shape, dtype=ACT_DTYPE, device=dev)
    v_out = torch.empty((B, g.Hout, g.Wout, g.Cout), dtype=torch.float32, device=dev) if want_v_out else None
    h_seq = torch.empty(out_shape, dtype=torch.float32, device=dev) if want_h else None
    _check_block_io(x, g, T, B, resid, v_in, decay, out_shape)
    d = _lib.BlockDesc(T=T, B=B, Hin=g.Hin, Win=g.Win, Cin=cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                       stride=g.stride, pad=g.pad, upsample=1 if g.kind == 'upconv' else 0, neuron=neuron, planes=planes,
                       gain=gain, v_th=v_th, v_reset=v_reset, tau=tau)
    if tsum is not None:
        assert tsum.dtype == ACT_DTYPE and tsum.is_contiguous() and tuple(tsum.shape) == (B, g.Hout, g.Wout, g.Cout)
    rc = _lib.lib().ss_conv_i8_fwd(ctypes.byref(d), _ptr(x), _ptr(w_i8), _ptr(wscale), _ptr(decay), _ptr(v_in), _ptr(v_out),
                                   _ptr(resid), _ptr(out), _ptr(h_seq), _ptr(tsum), _stream())
    _lib.check(rc, 'ss_conv_i8_fwd')
    return out, v_out, h_seq


def conv_neuron_fwd(x, geom, w_kn, *, T, B, in_layout, neuron, gain, v_th, v_reset, tau=2.0, decay=None,
                    v_in=None, want_v_out=False, resid=None, want_h=False):
    """fp32 CUDA-core fused block (ss_conv_neuron_fwd): exact fp32 weights; also reads the reference's fp32 NCHW frames.
    Returns (out u8 [T,B,Hout,Wout,Cout], v_out, h_seq)."""
    _require_cuda(x, 'x')
    dev = x.device
    g = geom
    if in_layout == SS_IN_U8_TBHWC:
        assert x.dtype == ACT_DTYPE and x.is_contiguous() and tuple(x.shape) == (T, B, g.Hin, g.Win, g.Cin), \
            (x.dtype, tuple(x.shape), (T, B, g.Hin, g.Win, g.Cin))
    else:
        assert x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, T, g.Cin, g.Hin, g.Win), \
            (x.dtype, tuple(x.shape), (B, T, g.Cin, g.Hin, g.Win))
    ymap, xmap = g.maps(dev)
    out_shape = (T, B, g.Hout, g.Wout, g.Cout)
    out = torch.empty(out_shape, dtype=ACT_DTYPE, device=dev)
    v_out = torch.empty((B, g.Hout, g.Wout, g.Cout), dtype=torch.float32, device=dev) if want_v_out else None
    h_seq = torch.empty(out_shape, dtype=torch.float32, device=dev) if want_h else None
    _check_block_io(x, g, T, B, resid, v_in, decay, out_shape)
    cg = _lib.ConvGeom(T=T, B=B, Hin=g.Hin, Win=g.Win, Cin=g.Cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                       in_layout=in_layout, neuron=neuron, reserved0=0, gain=gain, v_th=v_th, v_reset=v_reset,
                       tau=tau, reserved1=0, reserved2=0)
    rc = _lib.lib().ss_conv_neuron_fwd(ctypes.byref(cg), _ptr(x), _ptr(ymap), _ptr(xmap), _ptr(w_kn), _ptr(decay),
                                       _ptr(v_in), _ptr(v_out), _ptr(resid), _ptr(out), _ptr(h_seq), _stream())
    _lib.check(rc, 'ss_conv_neuron_fwd')
    return out, v_out, h_seq


def heads_fwd(acts, geoms, weights_9c, biases, *, T, B, H, W, gain, v_io, acts_sum=None):
    """Four prediction heads + I-neuron accumulation.  acts/geoms/weights/biases in execution order (head 4 first).
    ``acts_sum``: optional list of u8 [B,Hs,Ws,C] sums over the first T-1 timesteps -> 2 head passes instead of T.
    v_io fp32 [B,H,W] is updated in place.  Returns depths fp32 [4,B,H,W] (potential after each head, last step)."""
    dev = v_io.device
    a = _lib.HeadsArgs()
    a.T, a.B, a.H, a.W, a.gain = T, B, H, W, gain
    keep = []
    for i in range(4):
        g = geoms[i]
        assert acts[i].dtype == ACT_DTYPE and acts[i].is_contiguous() and \
            tuple(acts[i].shape) == (T, B, g.Hin, g.Win, g.Cin)
        ym, xm = g.maps(dev)
        ne = 2 if (acts_sum is not None and T > 1) else T
        tp = torch.empty((ne, B, 9, g.Hin, g.Win), dtype=torch.float32, device=dev)
        keep += [ym, xm, tp]
        a.taps[i] = tp.data_ptr()
        if acts_sum is not None and T > 1:
            assert acts_sum[i].dtype == ACT_DTYPE and acts_sum[i].is_contiguous() and \
                tuple(acts_sum[i].shape) == (B, g.Hin, g.Win, g.Cin)
            a.acts_sum[i] = acts_sum[i].data_ptr()
        a.C[i], a.Hs[i], a.Ws[i] = g.Cin, g.Hin, g.Win
        a.acts[i] = acts[i].data_ptr()
        a.w[i] = weights_9c[i].data_ptr()
        a.bias[i] = biases[i].data_ptr()
        a.ymap[i] = ym.data_ptr()
        a.xmap[i] = xm.data_ptr()
    depths = torch.empty((4, B, H, W), dtype=torch.float32, device=dev)
    rc = _lib.lib().ss_heads_fwd(ctypes.byref(a), _ptr(v_io), _ptr(depths), _stream())
    _lib.check(rc, 'ss_heads_fwd')
    return depths
