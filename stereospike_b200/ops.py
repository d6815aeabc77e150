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


# ----------------------------------------------------------------------------------------- folded upsampled conv
_FOLD_GROUPS = {0: ((0,), (1, 2), (3, 4)),      # class L: the output sits on the LAST copy of its source pixel  (pattern 0,1,1,2,2)
                1: ((0, 1), (2, 3), (4,))}      # class M: ... on the second-to-last copy                          (pattern 0,0,1,1,2)
_FOLD_PATTERNS = {0: (0, 1, 1, 2, 2), 1: (0, 0, 1, 1, 2)}


def fold_axis(n_in, n_out, ks=5):
    """Regular / irregular structure of  UpsamplingNearest2d(n_out+ks-1) -> valid conv(ks)  along one axis.
    Returns (omap int32 [2][n_in-2]: output coordinate of source position s for class L / M or -1, bands [(start, len)] of
    the output coordinates that do not follow either regular pattern)."""
    assert ks == 5
    n_up = n_out + ks - 1
    scale = np.float32(n_in) / np.float32(n_up)
    src = np.minimum(np.floor(np.arange(n_up, dtype=np.float32) * scale).astype(np.int64), n_in - 1)
    omap = np.full((2, n_in - 2), -1, dtype=np.int32)
    irregular = []
    for o in range(n_out):
        pat = tuple(int(src[o + k] - src[o]) for k in range(ks))
        s0 = int(src[o])
        hit = [c for c, pt in _FOLD_PATTERNS.items() if pt == pat]
        if hit and s0 < n_in - 2 and omap[hit[0], s0] < 0:
            omap[hit[0], s0] = o
        else:
            irregular.append(o)
    bands = []
    for o in irregular:
        if bands and o == bands[-1][0] + bands[-1][1]:
            bands[-1][1] += 1
        else:
            bands.append([o, 1])
    return omap, [tuple(b) for b in bands]


class FoldPlan:
    """Device tables of one folded NNConvUpsampling geometry (cached per (Hin, Win, Hout, Wout, device))."""

    def __init__(self, Hin, Win, Hout, Wout, device):
        ymap, ybands = fold_axis(Hin, Hout)
        xmap, xbands = fold_axis(Win, Wout)
        xb = []
        for st, ln in xbands:                      # column bands are at most one 8-wide tile each
            while ln > 0:
                xb.append((st, min(ln, 8)))
                st, ln = st + 8, ln - 8
        dev = torch.device(device)
        t = lambda a: torch.tensor(a, dtype=torch.int32, device=dev).contiguous()
        self.ymap, self.xmap = t(ymap), t(xmap)
        self.yband_start, self.yband_len = t([b[0] for b in ybands]), t([b[1] for b in ybands])
        self.xband_start, self.xband_len = t([b[0] for b in xb]), t([b[1] for b in xb])
        self.n_ybands, self.n_xbands = len(ybands), len(xb)
        self.yband_rows = max([b[1] for b in ybands]) if ybands else 0
        self.covered = float((ymap >= 0).sum()) / Hout * float((xmap >= 0).sum()) / Wout      # fraction of regular outputs


@functools.lru_cache(maxsize=None)
def fold_plan(Hin, Win, Hout, Wout, device):
    return FoldPlan(Hin, Win, Hout, Wout, device)


def pack_weights_folded(weight, planes):
    """Quantises a 5x5 NNConvUpsampling weight (OIHW) with 3 bits of head-room and returns
    (w_fold: digit planes of the four folded 3x3 weight sets, w_full: digit planes of the 5x5 taps (band passes),
    wscale fp32 [Cout]).  The folded sets are exact integer sums of the quantised taps, so both passes produce
    identical integers for every output they share."""
    w = weight.detach().double()
    co, ci, kh, kw = w.shape
    assert kh == 5 and kw == 5 and co % 32 == 0 and ci % 32 == 0
    m = w.abs().amax(dim=(1, 2, 3))
    ex = torch.where(m > 0, torch.floor(torch.log2(m)) + 1, torch.zeros_like(m))          # m < 2^ex
    e = ex - (8 * planes - 1) + 3
    q = torch.round(w * torch.pow(2.0, -e).view(-1, 1, 1, 1))
    assert float(q.abs().max()) < 2 ** (8 * planes - 4) + 1
    sets = []
    for cy in (0, 1):
        for cx in (0, 1):
            f = q.new_zeros(co, ci, 3, 3)
            for dy, kys in enumerate(_FOLD_GROUPS[cy]):
                for dx, kxs in enumerate(_FOLD_GROUPS[cx]):
                    for ky in kys:
                        for kx in kxs:
                            f[:, :, dy, dx] += q[:, :, ky, kx]
            sets.append(f)
    # weight-set index = output-channel tile * 4 + class  ->  stack as [tile][class][32] "output channels"
    fold = torch.stack(sets, dim=1).view(co // 32, 32, 4, ci, 3, 3).permute(0, 2, 1, 3, 4, 5).reshape(4 * co, ci, 3, 3)
    dev = weight.device
    zeros = torch.zeros(4 * co, dtype=torch.int32, device=dev)
    fold32 = fold.float().contiguous()
    full32 = q.float().contiguous()
    w_fold = torch.empty(4 * co * ci * 9 * planes, dtype=torch.int8, device=dev)
    w_full = torch.empty(co * ci * 25 * planes, dtype=torch.int8, device=dev)
    L = _lib.lib()
    _lib.check(L.ss_pack_digits_i8(_ptr(fold32), 4 * co, ci, 3, planes, _ptr(zeros), _ptr(w_fold), _stream()), 'ss_pack_digits_i8')
    _lib.check(L.ss_pack_digits_i8(_ptr(full32), co, ci, 5, planes, _ptr(zeros), _ptr(w_full), _stream()), 'ss_pack_digits_i8')
    wscale = torch.pow(2.0, e).float().contiguous()
    return w_fold, w_full, wscale


def _check_block_io(x, g, T, B, resid, v_in, decay, out_shape):
    if resid is not None:
        assert resid.dtype == ACT_DTYPE and resid.is_contiguous() and tuple(resid.shape) == out_shape
    if v_in is not None:
        assert v_in.dtype == torch.float32 and v_in.is_contiguous() and tuple(v_in.shape) == (B, g.Hout, g.Wout, g.Cout)
    if decay is not None:
        assert decay.dtype == torch.float32 and decay.numel() == 1 and decay.is_cuda


# ----------------------------------------------------------------------------------------- fused block
def conv_i8_fwd(x, geom, w_i8, wscale, *, T, B, neuron, gain, v_th, v_reset, tau=2.0, decay=None, v_in=None,
                want_v_out=False, resid=None, want_h=False, planes=3, cin=None, tsum=None, outputs=None, tile_maps=None,
                desc_override=None):
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
    if outputs is not None:
        out, v_out, h_seq = outputs           # a later pass of the same block writes into the first pass's tensors
    else:
        out = torch.empty(out_shape, dtype=ACT_DTYPE, device=dev)
        v_out = torch.empty((B, g.Hout, g.Wout, g.Cout), dtype=torch.float32, device=dev) if want_v_out else None
        h_seq = torch.empty(out_shape, dtype=torch.float32, device=dev) if want_h else None
    _check_block_io(x, g, T, B, resid, v_in, decay, out_shape)
    d = _lib.BlockDesc(T=T, B=B, Hin=g.Hin, Win=g.Win, Cin=cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                       stride=g.stride, pad=g.pad, upsample=1 if g.kind == 'upconv' else 0, neuron=neuron, planes=planes,
                       gain=gain, v_th=v_th, v_reset=v_reset, tau=tau)
    if desc_override:
        for k, v in desc_override.items():
            setattr(d, k, v)
    if tsum is not None:
        assert tsum.dtype == ACT_DTYPE and tsum.is_contiguous() and tuple(tsum.shape) == (B, g.Hout, g.Wout, g.Cout)
    if tile_maps is None:
        rc = _lib.lib().ss_conv_i8_fwd(ctypes.byref(d), _ptr(x), _ptr(w_i8), _ptr(wscale), _ptr(decay), _ptr(v_in), _ptr(v_out),
                                       _ptr(resid), _ptr(out), _ptr(h_seq), _ptr(tsum), _stream())
    else:
        rc = _lib.lib().ss_conv_i8_fwd_ex(ctypes.byref(d), ctypes.byref(tile_maps), _ptr(x), _ptr(w_i8), _ptr(wscale), _ptr(decay),
                                          _ptr(v_in), _ptr(v_out), _ptr(resid), _ptr(out), _ptr(h_seq), _ptr(tsum), _stream())
    _lib.check(rc, 'ss_conv_i8_fwd')
    return out, v_out, h_seq


def conv_i8_fwd_folded(x, geom, w_fold, w_full, wscale, **kw):
    """NNConvUpsampling block (5x5, ~2x) as four folded 3x3 convs on the source + the general kernel on the irregular
    row / column bands (three launches writing the same output tensors).  Same arguments / results as conv_i8_fwd."""
    g = geom
    assert g.kind == 'upconv' and g.ks == 5
    plan = fold_plan(g.Hin, g.Win, g.Hout, g.Wout, str(x.device))
    tm = _lib.TileMaps(mode=_lib.SS_TILES_FOLDED, nbands=0, band_rows=0, reserved=0, ymap_out=plan.ymap.data_ptr(),
                       xmap_out=plan.xmap.data_ptr(), band_start=0, band_len=0)
    res = conv_i8_fwd(x, g, w_fold, wscale, tile_maps=tm, desc_override=dict(ks=3, stride=1, pad=0, upsample=0), **kw)
    kw2 = dict(kw)
    kw2.pop('want_v_out', None)
    kw2.pop('want_h', None)
    if plan.n_ybands:
        tm = _lib.TileMaps(mode=_lib.SS_TILES_ROW_BANDS, nbands=plan.n_ybands, band_rows=plan.yband_rows, reserved=0, ymap_out=0,
                           xmap_out=0, band_start=plan.yband_start.data_ptr(), band_len=plan.yband_len.data_ptr())
        conv_i8_fwd(x, g, w_full, wscale, tile_maps=tm, outputs=res, **kw2)
    if plan.n_xbands:
        tm = _lib.TileMaps(mode=_lib.SS_TILES_COL_BANDS, nbands=plan.n_xbands, band_rows=8, reserved=0, ymap_out=0, xmap_out=0,
                           band_start=plan.xband_start.data_ptr(), band_len=plan.xband_len.data_ptr())
        conv_i8_fwd(x, g, w_full, wscale, tile_maps=tm, outputs=res, **kw2)
    return res


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


# ----------------------------------------------------------------------------------------- tensor-core gradients
def _pack_bf16(w_eff, ks, ntile):
    """OIHW fp32 [nsets*ntile][Cg][ks][ks] -> the correlation kernel's bf16 shared-memory image."""
    w_eff = w_eff.contiguous().float()
    co, ci = int(w_eff.shape[0]), int(w_eff.shape[1])
    img = torch.empty(co * ci * ks * ks, dtype=torch.bfloat16, device=w_eff.device)
    _lib.check(_lib.lib().ss_pack_weights_bf16(_ptr(w_eff), co, ci, ks, ntile, _ptr(img), _stream()), 'ss_pack_weights_bf16')
    return img


@functools.lru_cache(maxsize=None)
def _dgrad_tables(kind, stride, ks, Hin, Win, Hout, Wout, device):
    """Geometry-only device tables of a block's data gradient (cached: building them is a blocking host-to-device copy)."""
    dev = torch.device(device)
    t = lambda a: torch.tensor(a, dtype=torch.int32, device=dev).contiguous()
    if kind == 'conv' and stride == 1:
        return None, None, None
    if kind == 'conv':
        Hv, Wv = (Hin + 1) // 2, (Win + 1) // 2
        ym = np.full((2, Hv), -1, dtype=np.int32)
        xm = np.full((2, Wv), -1, dtype=np.int32)
        for par in (0, 1):
            yy = 2 * np.arange(Hv) + par
            xx = 2 * np.arange(Wv) + par
            ym[par] = np.where(yy < Hin, yy, -1)
            xm[par] = np.where(xx < Win, xx, -1)
        # filter tap read by correlation offset d of parity class par: 2*(2 - d) + par; 5 = the appended zero tap
        kidx = torch.tensor([[min(2 * (2 - d) + par, 5) for d in range(3)] for par in (0, 1)], device=dev)
        return t(ym), t(xm), kidx

    def src(n_in, n_up):
        scale = np.float32(n_in) / np.float32(n_up)
        return np.minimum(np.floor(np.arange(n_up, dtype=np.float32) * scale).astype(np.int64), n_in - 1).astype(np.int32)
    return t(src(Hin, Hout + ks - 1)), t(src(Win, Wout + ks - 1)), None


class DgradPlan:
    """Data gradient of one fused block as a single ss_corr_bf16 call (include/stereospike_b200.h): correlation weights
    derived from the block's OIHW weight, virtual grid, output maps.  ``geom`` is the block's forward BlockGeom."""

    def __init__(self, weight, geom, device):
        g = geom
        w = weight.detach().float()
        co, ci, ks, _ = w.shape
        self.ntile = 64 if ci % 64 == 0 else 32
        assert ci % 32 == 0 and co % 16 == 0
        self.ymap, self.xmap, kidx = _dgrad_tables(g.kind, g.stride, ks, g.Hin, g.Win, g.Hout, g.Wout, str(device))
        wt = w.flip(2, 3).transpose(0, 1)                       # [Cin][Cout][ks][ks]: correlation form of the transposed conv
        if g.kind == 'conv' and g.stride == 1:
            assert 2 * g.pad == ks - 1
            self.ks, self.pad, self.nclass, self.mode = ks, ks - 1 - g.pad, 1, _lib.SS_CORR_ACCUMULATE
            self.Hv, self.Wv = g.Hin, g.Win
            w_eff = wt
        elif g.kind == 'conv':
            assert g.stride == 2 and ks == 5 and g.pad == 2
            # input row y = 2i + py receives  sum_dy g[i - 1 + dy] * W[2*(2 - dy) + py]  (dy = 0..2; tap index > 4 -> absent)
            self.ks, self.pad, self.nclass, self.mode = 3, 1, 4, _lib.SS_CORR_ACCUMULATE
            self.Hv, self.Wv = (g.Hin + 1) // 2, (g.Win + 1) // 2
            wp = torch.nn.functional.pad(w, (0, 1, 0, 1))      # one gather instead of 36 slice copies
            f = wp[:, :, kidx[:, :, None, None], kidx[None, None, :, :]]          # [co][ci][py][dy][px][dx]
            nt = self.ntile
            # -> [ci-tile][class = 2*py + px][nt][co][dy][dx]
            w_eff = f.permute(1, 2, 4, 0, 3, 5).reshape(ci // nt, nt, 4, co, 3, 3).permute(0, 2, 1, 3, 4, 5).reshape(4 * ci, co, 3, 3)
        else:
            # NNConvUpsampling: gradient w.r.t. the virtual upsampled image, routed to its nearest-neighbour source pixel
            self.ks, self.pad, self.nclass, self.mode = ks, ks - 1, 1, _lib.SS_CORR_ATOMIC
            self.Hv, self.Wv = g.Hout + ks - 1, g.Wout + ks - 1
            w_eff = wt
        self.w_img = _pack_bf16(w_eff, self.ks, self.ntile)
        self.geom = g

    def run(self, g_bf16, dst, T, B):
        g = self.geom
        assert g_bf16.dtype == torch.bfloat16 and g_bf16.is_contiguous() and tuple(g_bf16.shape) == (T, B, g.Hout, g.Wout, g.Cout)
        assert dst.dtype == torch.float32 and dst.is_contiguous() and tuple(dst.shape) == (T, B, g.Hin, g.Win, g.Cin)
        d = _lib.CorrDesc(T=T, B=B, Hg=g.Hout, Wg=g.Wout, Cg=g.Cout, Hv=self.Hv, Wv=self.Wv, Hdst=g.Hin, Wdst=g.Win, Cdst=g.Cin,
                          ks=self.ks, pad=self.pad, nclass=self.nclass, out_mode=self.mode, ntile=self.ntile, reserved=0)
        _lib.check(_lib.lib().ss_corr_bf16(ctypes.byref(d), _ptr(g_bf16), _ptr(self.w_img), _ptr(self.ymap), _ptr(self.xmap),
                                           _ptr(dst), _stream()), 'ss_corr_bf16')


def conv_wgrad_bf16(x, g_bf16, geom, T, B, cin=None):
    """Weight gradient of one fused block on the tensor cores -> fp32 [K][Cout] (k = (ky*ks + kx)*Cin + c)."""
    g = geom
    cin = g.Cin if cin is None else cin
    assert x.dtype == ACT_DTYPE and x.is_contiguous() and tuple(x.shape) == (T, B, g.Hin, g.Win, cin)
    assert g_bf16.dtype == torch.bfloat16 and g_bf16.is_contiguous() and tuple(g_bf16.shape) == (T, B, g.Hout, g.Wout, g.Cout)
    g_w = torch.zeros((g.ks * g.ks * cin, g.Cout), dtype=torch.float32, device=x.device)
    d = _lib.BlockDesc(T=T, B=B, Hin=g.Hin, Win=g.Win, Cin=cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                       stride=g.stride, pad=g.pad, upsample=1 if g.kind == 'upconv' else 0, neuron=0, planes=3,
                       gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0)
    _lib.check(_lib.lib().ss_conv_wgrad_bf16(ctypes.byref(d), _ptr(x), _ptr(g_bf16), _ptr(g_w), _stream()), 'ss_conv_wgrad_bf16')
    return g_w
