"""Tensor-level wrappers over the C ABI: geometry tables, weight packing, fused block / heads calls.

PyTorch is used here for device memory, streams and autograd plumbing only; all arithmetic of the hot path
happens inside libstereospike_b200.so.
"""
import ctypes
import os
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


def first_layer_channels(C):
    """Channels of the packed first-layer input for C frame channels: 4 (im2col mode) up to 4, else the next multiple of 32."""
    return 4 if C <= 4 else (C + 31) // 32 * 32


def pack_events(x_seq, status=None):
    """fp32 [B,T,C,H,W] event-count frames -> u8 [T,B,H,W,Cpad] (the first block's tensor-core input; Cpad = 4, or a multiple of
    32 for the channel-concatenated temporal mode with C > 4 channels)."""
    _require_cuda(x_seq, 'x')
    B, T, C, H, W = x_seq.shape
    cpad = first_layer_channels(C)
    out = torch.empty((T, B, H, W, cpad), dtype=torch.uint8, device=x_seq.device)
    _lib.check(_lib.lib().ss_pack_events_c(_ptr(x_seq), B, T, C, cpad, H, W, _ptr(out), _ptr(status), _stream()), 'ss_pack_events')
    return out


# ----------------------------------------------------------------------------------------- folded upsampled conv
FOLD_OVERLAP = 0 if os.environ.get('SS_FOLD_OVERLAP', '1') == '0' else 1   # the three passes of a folded block overlap their tails
FOLD_ROWS_BY_LIST = os.environ.get('SS_FOLD_ROWS_BY_LIST', '1') != '0'    # dense folded pass on listed regular rows (small decoder blocks)

# Replication patterns of the 5 taps of one axis of  UpsamplingNearest2d(n_out + 4) -> valid conv(5)  (source index of tap k
# relative to the first one): the two regular ones and the three next to a 3-fold replication (include/stereospike_b200.h).
FOLD_PATTERNS = ((0, 1, 1, 2, 2), (0, 0, 1, 1, 2), (0, 0, 0, 1, 1), (0, 1, 1, 1, 2), (0, 0, 1, 1, 1))     # L, M, A, B, C


def _fold_groups(pattern):
    """filter taps that read source offset d = 0, 1, 2."""
    return tuple(tuple(k for k in range(5) if pattern[k] == d) for d in range(3))


def fold_axis(n_in, n_out, ks=5):
    """Per output coordinate o of one axis: first source pixel s0[o] and replication pattern pat[o] (index into FOLD_PATTERNS,
    -1 = none of them: the geometry cannot be folded)."""
    assert ks == 5
    n_up = n_out + ks - 1
    scale = np.float32(n_in) / np.float32(n_up)
    src = np.minimum(np.floor(np.arange(n_up, dtype=np.float32) * scale).astype(np.int64), n_in - 1)
    s0 = src[:n_out].astype(np.int64)
    pat = np.full(n_out, -1, dtype=np.int64)
    for o in range(n_out):
        rel = tuple(int(src[o + k] - src[o]) for k in range(ks))
        if rel in FOLD_PATTERNS:
            pat[o] = FOLD_PATTERNS.index(rel)
    return s0, pat


class FoldPlan:
    """Device tables of one folded NNConvUpsampling geometry and batch size (cached).  Three passes write the block's output:
    dense (4 regular class pairs, 3x3 on the source), irregular rows x all columns, irregular columns x regular rows."""

    def __init__(self, Hin, Win, Hout, Wout, B, device):
        s0y, py = fold_axis(Hin, Hout)
        s0x, px = fold_axis(Win, Wout)
        dev = torch.device(device)
        t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32, device=dev).contiguous()
        self.ok = bool((py >= 0).all() and (px >= 0).all() and (s0y + 2 < Hin).all() and (s0x + 2 < Win).all() and
                       Hin >= 3 and Win >= 3 and (B * max(Hin * Win, Hout * Wout)) < 2 ** 31)
        if not self.ok:
            return

        def dense_map(s0, pat, n_in):
            m = np.full((2, n_in - 2), -1, dtype=np.int32)
            for o in range(len(s0)):
                if pat[o] < 2:
                    assert m[pat[o], s0[o]] < 0
                    m[pat[o], s0[o]] = o
            return m
        self.ymap, self.xmap = t32(dense_map(s0y, py, Hin)), t32(dense_map(s0x, px, Win))

        def lists(s0, pat, src_off, out_off, classes=(2, 3, 4)):
            """[len(classes)][n] source / output offsets of the coordinates of every sample whose replication pattern is in
            ``classes`` (default: the three irregular ones), padded with -1."""
            per = [[o for o in range(len(s0)) if pat[o] == c] for c in classes]
            n = max(1, B * max(len(v) for v in per))
            src = np.full((len(classes), n), -1, dtype=np.int64)
            out = np.full((len(classes), n), -1, dtype=np.int64)
            for c in range(len(classes)):
                i = 0
                for b in range(B):
                    for o in per[c]:
                        src[c, i], out[c, i] = src_off(b, s0[o]), out_off(b, o)
                        i += 1
            return t32(src), t32(out), n, sum(len(v) for v in per)

        def concat_lists(s0, pat, src_off, out_off, classes):
            """The same lists CONCATENATED, each class padded to whole 16-entry tiles on its own (ss_tile_maps.item_tab):
            (src, out, total entries, [(first tile row, tile rows)] per class)."""
            srcs, outs, ranges, ty = [], [], [], 0
            for c in classes:
                ent = [(src_off(b, s0[o]), out_off(b, o)) for b in range(B) for o in range(len(s0)) if pat[o] == c]
                nt = (len(ent) + 15) // 16
                ent += [(-1, -1)] * (nt * 16 - len(ent))
                srcs += [e[0] for e in ent]
                outs += [e[1] for e in ent]
                ranges.append((ty, nt))
                ty += nt
            if ty == 0:
                srcs, outs = [-1] * 16, [-1] * 16
            return t32(np.asarray(srcs, dtype=np.int64)), t32(np.asarray(outs, dtype=np.int64)), len(srcs), tuple(ranges)
        self._t32 = t32
        self.crow = concat_lists(s0y, py, lambda b, s: (b * Hin + s) * Win, lambda b, o: (b * Hout + o) * Wout, (2, 3, 4))
        self.ccol = concat_lists(s0x, px, lambda b, s: b * Hin * Win + s, lambda b, o: b * Hout * Wout + o, (2, 3, 4))
        self.creg = concat_lists(s0y, py, lambda b, s: (b * Hin + s) * Win, lambda b, o: (b * Hout + o) * Wout, (0, 1))
        self._items = {}
        self.row_src, self.row_out, self.row_n, self.n_irr_rows = lists(
            s0y, py, lambda b, s: (b * Hin + s) * Win, lambda b, o: (b * Hout + o) * Wout)
        self.col_src, self.col_out, self.col_n, self.n_irr_cols = lists(
            s0x, px, lambda b, s: b * Hin * Win + s, lambda b, o: b * Hout * Wout + o)
        self.row_regular = torch.as_tensor((py < 2).astype(np.uint8), device=dev).contiguous()       # rows the column pass owns
        n_reg_r, n_reg_c = int((py < 2).sum()), int((px < 2).sum())
        # the regular rows of each row class as a list: the dense pass restricted to them (ss_tile_maps.transposed = 2).  The plain
        # dense pass evaluates every source row for both row classes and discards the irregular ones; worth replacing when a
        # sizeable part of the rows is irregular (the 17 -> 33 and 33 -> 65 blocks: 45 % / 23 %) -- the list form pays 3 private
        # source rows per tile row in patch traffic, which the weight-streaming blocks do not notice
        self.reg_src, self.reg_out, self.reg_n, _ = lists(s0y, py, lambda b, s: (b * Hin + s) * Win, lambda b, o: (b * Hout + o) * Wout,
                                                          classes=(0, 1))
        self.rows_by_list = n_reg_r < 0.85 * Hout
        self.covered = n_reg_r * n_reg_c / float(Hout * Wout)                 # fraction of outputs in the dense (9-tap) pass
        # taps executed per output, averaged (25 = unfolded): what bench.py credits a folded block with
        self.taps_per_output = (9.0 * n_reg_r * n_reg_c + 15.0 * self.n_irr_rows * Wout + 15.0 * self.n_irr_cols * n_reg_r) / float(Hout * Wout)


    def item_table(self, which, ntiles_out, tiles_x, col_classes=1):
        """Device int32 [n_items][2] = {weight set, m-tile} of a row-list pass over the concatenated lists ``which`` ('crow', 'ccol',
        'creg'): per output-channel tile, per (row class, column class), the tiles of that row class -- a class gets exactly the
        tiles it needs.  Cached."""
        key = (which, ntiles_out, tiles_x, col_classes)
        if key not in self._items:
            ranges = getattr(self, which)[3]
            nclass = len(ranges) * col_classes
            items = []
            for n in range(ntiles_out):
                for rc, (ty0, nt) in enumerate(ranges):
                    for cc in range(col_classes):
                        wset = n * nclass + rc * col_classes + cc
                        for ty in range(ty0, ty0 + nt):
                            items += [(wset, ty * tiles_x + tx) for tx in range(tiles_x)]
            tab = self._t32(np.asarray(items, dtype=np.int64).reshape(-1, 2)) if items else None
            self._items[key] = (tab, len(items))
        return self._items[key]


@functools.lru_cache(maxsize=None)
def fold_plan(Hin, Win, Hout, Wout, B, device):
    return FoldPlan(Hin, Win, Hout, Wout, B, device)


def fold_weight_sets(weight, planes):
    """Quantises a 5x5 NNConvUpsampling weight (OIHW) with head-room for the tap sums.  Returns (q: the quantised taps,
    dense: the four regular 3x3 sets [class = 2*row_class + col_class], rows: the three 3x5 sets (dy, kx) of the irregular rows,
    cols: the three 3x5 sets (dx, ky) of the irregular columns in the transposed frame, e: exponent per output channel), all
    float64 tensors holding exact integers.  Every set is an exact sum of the same quantised taps, so every pass produces the
    integers of the 25-tap conv with those taps."""
    w = weight.detach().double()
    co, ci, kh, kw = w.shape
    assert kh == 5 and kw == 5 and co % 32 == 0 and ci % 32 == 0
    m = w.abs().amax(dim=(1, 2, 3))
    ex = torch.where(m > 0, torch.floor(torch.log2(m)) + 1, torch.zeros_like(m))          # m < 2^ex
    groups = [_fold_groups(pt) for pt in FOLD_PATTERNS]
    limit = 127.0 * 256.0 ** (planes - 1)               # top balanced digit <= 127
    # head-room: a folded weight sums at most 4 taps (2 x 2 in the dense sets, 3 in the row / column sets), i.e. 2 bits in the worst
    # case and usually 1-2; start with 1 bit and give one more bit only to the output channels whose sums overflow
    e = ex - (8 * planes - 1) + 1
    for _ in range(4):
        q = torch.round(w * torch.pow(2.0, -e).view(-1, 1, 1, 1))

        def fold_rows(t, g):        # t [co,ci,5,n] -> [co,ci,3,n]: sum the filter rows of each group
            return torch.stack([t[:, :, list(k), :].sum(2) if k else t.new_zeros(co, ci, t.shape[3]) for k in g], dim=2)
        dense = []
        for cy in (0, 1):
            for cx in (0, 1):
                f = fold_rows(q, groups[cy])                                              # [co,ci,3,5]
                f = fold_rows(f.transpose(2, 3), groups[cx]).transpose(2, 3)              # [co,ci,3,3]
                dense.append(f)
        rows = [fold_rows(q, groups[2 + c]) for c in range(3)]                            # [co,ci,3(dy),5(kx)]
        cols = [fold_rows(q.transpose(2, 3), groups[2 + c]) for c in range(3)]            # [co,ci,3(dx),5(ky)]
        over = torch.stack([torch.stack(v).abs().amax(dim=(0, 2, 3, 4)) for v in (dense, rows, cols)]).amax(0) > limit
        if not bool(over.any()):
            break
        e = e + over.to(e.dtype)        # one more bit of head-room for the output channels whose sums do not fit
    else:
        raise RuntimeError('pack_weights_folded: folded tap sums do not fit the digit planes')
    return q, dense, rows, cols, e


def pack_weights_folded(weight, planes, with_exp=False):
    """Digit-plane images (dense, rows, cols) of a folded NNConvUpsampling block + wscale fp32 [Cout]: one launch of
    ss_pack_weights_folded (bit-identical to pack_weights_folded_host below, which derives them with torch ops)."""
    _require_cuda(weight, 'weight')
    w = weight.detach().contiguous().float()
    co, ci, kh, kw = (int(v) for v in w.shape)
    assert kh == 5 and kw == 5 and co % 32 == 0 and ci % 32 == 0
    dev = w.device
    dense = torch.empty(4 * co * ci * 9 * planes, dtype=torch.int8, device=dev)
    rows = torch.empty(3 * co * ci * 15 * planes, dtype=torch.int8, device=dev)
    cols = torch.empty(3 * co * ci * 15 * planes, dtype=torch.int8, device=dev)
    wscale = torch.empty(co, dtype=torch.float32, device=dev)
    wexp = torch.empty(co, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ss_pack_weights_folded(_ptr(w), co, ci, planes, _ptr(dense), _ptr(rows), _ptr(cols), _ptr(wscale), _ptr(wexp),
                                                 _stream()), 'ss_pack_weights_folded')
    if with_exp:
        return dense, rows, cols, wscale, wexp
    return dense, rows, cols, wscale


def pack_weights_with_exponents(weight, planes, wexp):
    """Plain (unfolded) digit-plane image of ``weight`` quantised with the given per-output-channel exponents -- with the fold's
    exponents these are the SAME quantised taps the folded sets are sums of: what a call too small to be worth folding runs, so
    that a sample's result does not depend on the size of the batch it arrives in."""
    w = weight.detach().contiguous().float()
    co, ci, kh, kw = (int(v) for v in w.shape)
    img = torch.empty(co * ci * kh * kw * planes, dtype=torch.int8, device=w.device)
    _lib.check(_lib.lib().ss_pack_digits_i8(_ptr(w), co, ci, kh, planes, _ptr(wexp), _ptr(img), _stream()), 'ss_pack_digits_i8')
    return img


def pack_weights_folded_host(weight, planes):
    """The same images derived on the host side of the C ABI (fold_weight_sets with torch ops, then ss_pack_digits_i8): the
    executable specification of ss_pack_weights_folded (tests/test_gpu_parity.py::test_fold_pack_kernel_matches_host_derivation)."""
    _, dense, rows, cols, e = fold_weight_sets(weight, planes)
    co, ci = int(weight.shape[0]), int(weight.shape[1])

    def stack_sets(sets):
        """weight-set index = output-channel tile * nclass + class  ->  [tile][class][32] "output channels"."""
        n = len(sets)
        t = torch.stack(sets, dim=1)                                                     # [co, n, ci, kh, kw]
        return t.view(co // 32, 32, n, *t.shape[2:]).permute(0, 2, 1, 3, 4, 5).reshape(n * co, *t.shape[2:]).float().contiguous()
    pack = lambda t: pack_digits_i8(t.to(weight.device), planes)
    return pack(stack_sets(dense)), pack(stack_sets(rows)), pack(stack_sets(cols)), torch.pow(2.0, e).float().contiguous().to(weight.device)


def pack_digits_i8(q, planes):
    """Digit-plane image of ALREADY QUANTISED integer weights (OIHW, exact integers in floating point; rectangular filters
    allowed): ss_pack_digits_i8_rect."""
    q = q.float().contiguous()
    _require_cuda(q, 'q')
    cop, ci, ky, kx = q.shape
    zeros = torch.zeros(cop, dtype=torch.int32, device=q.device)
    img = torch.empty(cop * ci * ky * kx * planes, dtype=torch.int8, device=q.device)
    _lib.check(_lib.lib().ss_pack_digits_i8_rect(_ptr(q), cop, ci, ky, kx, planes, _ptr(zeros), _ptr(img), _stream()),
               'ss_pack_digits_i8_rect')
    return img


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
                desc_override=None, stats=None, independent_steps=False):
    """Tensor-core fused block over all T timesteps (ss_conv_i8_fwd).  x: u8 [T,B,Hin,Win,Cin].
    ``independent_steps``: the T steps are independent samples (stateless inference; ss_tile_maps.independent_steps).
    ``tsum``: optional u8 [B,Hout,Wout,Cout] receiving the sum of the first T-1 output steps (input of the linear heads).
    ``stats``: optional int64 [6] device tensor the launch ADDS its firing statistics to ({spikes, nonzero outputs, sum out^2} over
    all steps, then over the last step).
    Returns (out u8 [T,B,Hout,Wout,Cout], v_out, h_seq)."""
    _require_cuda(x, 'x')
    dev = x.device
    g = geom
    cin = g.Cin if cin is None else cin
    if stats is not None:
        assert stats.dtype == torch.int64 and stats.numel() == 6 and stats.is_contiguous() and stats.is_cuda
        if tile_maps is None:
            tile_maps = _lib.TileMaps(mode=_lib.SS_TILES_PLAIN, nclass=1, rl_n=0, transposed=0, ymap_out=0, xmap_out=0, rl_src=0,
                                      rl_out=0, rl_collive=0, stats=0)
        tile_maps.stats = stats.data_ptr()
    if independent_steps:
        if tile_maps is None:
            tile_maps = _lib.TileMaps(mode=_lib.SS_TILES_PLAIN, nclass=1, rl_n=0, transposed=0, ymap_out=0, xmap_out=0, rl_src=0,
                                      rl_out=0, rl_collive=0, stats=0)
        tile_maps.independent_steps = 1
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


def conv_i8_fwd_folded(x, geom, w_dense, w_rows, w_cols, wscale, **kw):
    """NNConvUpsampling block (5x5, ~2x) as: four folded 3x3 convs on the source for the regular outputs, then the irregular
    rows and the irregular columns (3 + 2 output rows / columns per 3-fold replication) with one axis folded -- three
    launches writing the same output tensors.  Same arguments / results as conv_i8_fwd."""
    g = geom
    assert g.kind == 'upconv' and g.ks == 5
    plan = fold_plan(g.Hin, g.Win, g.Hout, g.Wout, int(kw['B']), str(x.device))
    assert plan.ok, 'geometry cannot be folded'
    nto = g.Cout // 32
    if plan.rows_by_list and g.Cin % 64 == 0 and FOLD_ROWS_BY_LIST:
        # regular rows x regular columns: the dense 3x3 sets on the listed regular rows only
        src, out, n, _ = plan.creg
        tab, nit = plan.item_table('creg', nto, (g.Win - 2 + 7) // 8, col_classes=2)
        tm = _lib.TileMaps(mode=_lib.SS_TILES_ROW_LIST, nclass=4, rl_n=n, transposed=2, ymap_out=0, xmap_out=plan.xmap.data_ptr(),
                           rl_src=src.data_ptr(), rl_out=out.data_ptr(), rl_collive=0, stats=0, item_tab=tab.data_ptr(), n_items=nit)
        res = conv_i8_fwd(x, g, w_dense, wscale, tile_maps=tm, **kw)
    else:
        tm = _lib.TileMaps(mode=_lib.SS_TILES_FOLDED, nclass=4, rl_n=0, transposed=0, ymap_out=plan.ymap.data_ptr(),
                           xmap_out=plan.xmap.data_ptr(), rl_src=0, rl_out=0, rl_collive=0, stats=0)
        res = conv_i8_fwd(x, g, w_dense, wscale, tile_maps=tm, desc_override=dict(ks=3, stride=1, pad=0, upsample=0), **kw)
    kw2 = dict(kw)
    for k in ('want_v_out', 'want_h', 'outputs'):
        kw2.pop(k, None)
    # the irregular rows / columns: class lists concatenated, every class with its own number of tiles (item table)
    if plan.n_irr_rows:
        src, out, n, _ = plan.crow
        tab, nit = plan.item_table('crow', nto, (g.Wout + 7) // 8)
        tm = _lib.TileMaps(mode=_lib.SS_TILES_ROW_LIST, nclass=3, rl_n=n, transposed=0, ymap_out=0, xmap_out=0, rl_src=src.data_ptr(),
                           rl_out=out.data_ptr(), rl_collive=0, stats=0, item_tab=tab.data_ptr(), n_items=nit, defer_wait=FOLD_OVERLAP)
        conv_i8_fwd(x, g, w_rows, wscale, tile_maps=tm, outputs=res, **kw2)
    if plan.n_irr_cols:
        src, out, n, _ = plan.ccol
        tab, nit = plan.item_table('ccol', nto, (g.Hout + 7) // 8)
        tm = _lib.TileMaps(mode=_lib.SS_TILES_ROW_LIST, nclass=3, rl_n=n, transposed=1, ymap_out=0, xmap_out=0, rl_src=src.data_ptr(),
                           rl_out=out.data_ptr(), rl_collive=plan.row_regular.data_ptr(), stats=0, item_tab=tab.data_ptr(), n_items=nit,
                           defer_wait=FOLD_OVERLAP)
        conv_i8_fwd(x, g, w_cols, wscale, tile_maps=tm, outputs=res, **kw2)
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


def conv_wgrad_bf16(x, g_bf16, geom, T, B, cin=None, x_full_range=False):
    """Weight gradient of one fused block on the tensor cores -> fp32 [K][Cout] (k = (ky*ks + kx)*Cin + c).
    x_full_range: x may hold any u8 value (event-count frames); otherwise every value must be < 128 (spikes, spike sums)."""
    g = geom
    cin = g.Cin if cin is None else cin
    assert x.dtype == ACT_DTYPE and x.is_contiguous() and tuple(x.shape) == (T, B, g.Hin, g.Win, cin)
    assert g_bf16.dtype == torch.bfloat16 and g_bf16.is_contiguous() and tuple(g_bf16.shape) == (T, B, g.Hout, g.Wout, g.Cout)
    g_w = torch.zeros((g.ks * g.ks * cin, g.Cout), dtype=torch.float32, device=x.device)
    d = _lib.BlockDesc(T=T, B=B, Hin=g.Hin, Win=g.Win, Cin=cin, Hout=g.Hout, Wout=g.Wout, Cout=g.Cout, ks=g.ks,
                       stride=g.stride, pad=g.pad, upsample=1 if g.kind == 'upconv' else 0, neuron=0, planes=0 if x_full_range else 3,
                       gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0)
    _lib.check(_lib.lib().ss_conv_wgrad_bf16(ctypes.byref(d), _ptr(x), _ptr(g_bf16), _ptr(g_w), _stream()), 'ss_conv_wgrad_bf16')
    return g_w
