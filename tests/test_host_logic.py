"""CPU: host-side logic of the drop-in modules -- gather tables, weight packing, state-dict compatibility, helpers."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import stereospike_b200 as sb
from stereospike_b200 import ops
from oracle import ref_model as rm
from oracle import sj_compat as sj

PAIRS = [(17, 33), (22, 44), (33, 65), (44, 87), (65, 130), (87, 173), (130, 260), (173, 346)]


@pytest.mark.parametrize('n_in,n_out', PAIRS + [(33, 260), (44, 346), (260, 260), (346, 346)])
@pytest.mark.parametrize('ks', [3, 5])
def test_upsample_map_matches_aten(n_in, n_out, ks):
    """The gather table must reproduce torch's nearest-neighbour index rule for every size pair the model uses."""
    n_up = n_out + ks - 1
    src = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1)
    up = F.interpolate(src, size=(n_up, 1), mode='nearest').view(-1).long().numpy()
    m = ops.upsample_axis_map(n_in, n_out, ks)
    assert m.shape == (n_out, ks)
    for k in range(ks):
        assert np.array_equal(m[:, k], up[k:k + n_out])


def test_conv_map_equals_unfold():
    m = ops.conv_axis_map(11, ops.conv_out_size(11, 5, 2, 2), 5, 2, 2)
    assert m.shape == (6, 5)
    assert list(m[0]) == [-1, -1, 0, 1, 2] and list(m[5]) == [8, 9, 10, -1, -1]


def test_maps_reproduce_conv_and_upconv():
    """conv via (ymap,xmap) gather + [K][Cout] weights == F.conv2d, for both block kinds (pure torch, CPU)."""
    torch.manual_seed(0)
    for kind in ('conv', 'upconv'):
        Cin, Cout, ks = 8, 32, 5
        x = torch.randn(2, Cin, 9, 11)
        w = torch.randn(Cout, Cin, ks, ks)
        if kind == 'conv':
            ref = F.conv2d(x, w, stride=2, padding=2)
            Ho, Wo = ref.shape[2:]
            ym, xm = ops.conv_axis_map(9, Ho, ks, 2, 2), ops.conv_axis_map(11, Wo, ks, 2, 2)
        else:
            Ho, Wo = 19, 23
            ref = F.conv2d(F.interpolate(x, size=(Ho + ks - 1, Wo + ks - 1), mode='nearest'), w)
            ym, xm = ops.upsample_axis_map(9, Ho, ks), ops.upsample_axis_map(11, Wo, ks)
        w_kn = ops.weight_to_kn(w)
        xp = torch.cat([x, torch.zeros(2, Cin, 1, 11)], 2)
        xp = torch.cat([xp, torch.zeros(2, Cin, 10, 1)], 3)       # index -1 -> the zero row / column
        cols = []
        for ky in range(ks):
            for kx in range(ks):
                g = xp[:, :, torch.from_numpy(ym[:, ky]).long()][:, :, :, torch.from_numpy(xm[:, kx]).long()]
                cols.append(g)                                   # [B,Cin,Ho,Wo]
        A = torch.stack(cols, 1).permute(0, 3, 4, 1, 2).reshape(2 * Ho * Wo, ks * ks * Cin)
        got = (A @ w_kn).reshape(2, Ho, Wo, Cout).permute(0, 3, 1, 2)
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
        assert torch.equal(ops.kn_to_weight(w_kn, Cout, Cin, ks), w)


@pytest.mark.parametrize('cls,kw,variant,mono', [
    (sb.StereoSpike, dict(multiply_factor=5.0), 'if', False),
    (sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike, dict(use_plif=True, tau=3.0), 'plif', False),
    (sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike, dict(use_plif=False, tau=3.0), 'lif', False),
    (sb.fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike, dict(use_plif=False, tau=3.0), 'lif', True)])
def test_state_dict_is_interchangeable_with_reference_layout(cls, kw, variant, mono):
    net = cls(**kw)
    oracle = rm.SpikingUNet(variant, mono, tau=3.0)
    sd_o, sd_n = oracle.state_dict(), net.state_dict()
    assert list(sd_o) == list(sd_n)
    for k in sd_o:
        assert sd_o[k].shape == sd_n[k].shape, k
    net.load_state_dict(sd_o)
    oracle.load_state_dict(net.state_dict())
    assert net.count_trainable_params() == sum(p.numel() for p in oracle.parameters())


def test_constructor_quirks_match_reference():
    # SNN_models.py:71-72: StereoSpike never forwards v_threshold / v_reset
    n = sb.StereoSpike(surrogate_function=sb.surrogate.ATan(), v_threshold=0.5, v_reset=0.3)
    assert n.bottom[2].v_threshold == 1.0 and n.bottom[2].v_reset == 0.0
    # blocks.py:142: bottleneck neurons silently use the default Sigmoid surrogate
    assert isinstance(n.bottleneck[0].sn1.surrogate_function, sb.surrogate.Sigmoid)
    assert isinstance(n.bottom[2].surrogate_function, sb.surrogate.ATan)
    assert n.Ineurons.v_threshold == float('inf')
    m = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0)
    assert isinstance(m.conv1[2], sb.neuron.LIFNode) and isinstance(m.bottleneck[1].sn2, sb.neuron.ParametricLIFNode)
    assert float(m.bottleneck[0].sn1.w) == pytest.approx(float(sj.ParametricLIFNode(3.0).w))


def test_reset_and_state_helpers():
    n = sb.StereoSpike()
    nodes = [m for m in n.modules() if hasattr(m, 'reset')]
    assert len(nodes) == 14                       # 13 spiking layers + the I-neuron pool
    n.conv1[2].v = torch.ones(1, 64, 130, 173)
    n.set_init_depths_potentials(torch.full((1, 1, 260, 346), 2.0))
    assert len(n.get_network_state()) == 14
    sb.functional.reset_net(n)
    assert n.conv1[2].v == 0.0 and n.Ineurons.v == 0.0
    n.increment_epoch()
    n.update_max_accuracy(0.5)
    assert n.epoch == 1 and n.get_max_accuracy() == 0.5


def test_engine_wiring():
    e = sb.StereoSpike().engine
    assert [s.name for s in e.sites] == ['bottom', 'conv1', 'conv2', 'conv3', 'conv4', 'bottleneck.0.conv1',
                                         'bottleneck.0.conv2', 'bottleneck.1.conv1', 'bottleneck.1.conv2', 'deconv4',
                                         'deconv3', 'deconv2', 'deconv1']
    assert [(s.src, s.resid) for s in e.sites[-4:]] == [('out_rconv', 'out_conv3'), ('out_add4', 'out_conv2'),
                                                        ('out_add3', 'out_conv1'), ('out_add2', 'out_bottom')]
    g = e.sites[-1].geom(130, 173)
    assert (g.Hout, g.Wout, g.K) == (260, 346, 1600)
    assert [h.src for h in e.heads] == ['out_add4', 'out_add3', 'out_add2', 'out_add1']


@pytest.mark.parametrize('Hin,Win,Hout,Wout', [(17, 22, 33, 44), (33, 44, 65, 87), (65, 87, 130, 173), (130, 173, 260, 346),      # deconv4..1
                                                (9, 11, 17, 21), (20, 27, 40, 53), (37, 13, 75, 26), (5, 10, 7, 18)])
def test_fold_plan_and_weight_sets_reproduce_upsampled_conv(Hin, Win, Hout, Wout):
    """Host logic of the folded NNConvUpsampling block, emulated in float64 on CPU: the dense 3x3 pass on the source (4 class
    pairs, output maps), the irregular-row pass (row lists, rows folded to 3 taps, 5 taps over the upsampled columns) and the
    irregular-column pass (transposed) together write EVERY output exactly once (the column pass masks the rows of the row
    pass) and equal UpsamplingNearest2d -> conv5x5 with the same quantised taps, integer for integer."""
    B, Cin, Cout = 2, 32, 32
    plan = ops.FoldPlan(Hin, Win, Hout, Wout, B, 'cpu')
    assert plan.ok
    g = torch.Generator().manual_seed(Hin)
    w = (torch.rand(Cout, Cin, 5, 5, generator=g) * 2 - 1) / 20
    x = (torch.rand(B, Cin, Hin, Win, generator=g) < 0.3).double() * torch.randint(1, 4, (B, Cin, Hin, Win), generator=g).double()
    q, dense, rows, cols, e = ops.fold_weight_sets(w, 3)
    assert float((q * torch.pow(2.0, e).view(-1, 1, 1, 1) - w.double()).abs().max()) < 2.0 ** -18
    ref = F.conv2d(F.interpolate(x, size=(Hout + 4, Wout + 4), mode='nearest'), q)         # [B,Cout,Hout,Wout] exact integers
    out = torch.full_like(ref, float('nan'))
    writes = torch.zeros(B, Hout, Wout)
    ymap, xmap = plan.ymap.numpy(), plan.xmap.numpy()
    # dense pass: virtual 3x3 pad-0 conv on the source per class pair
    for cy in (0, 1):
        for cx in (0, 1):
            v = F.conv2d(x, dense[2 * cy + cx])                                             # [B,Cout,Hin-2,Win-2]
            for sy in range(Hin - 2):
                oy = ymap[cy, sy]
                if oy < 0:
                    continue
                sel = xmap[cx] >= 0
                out[:, :, oy, xmap[cx][sel]] = v[:, :, sy, sel]
                writes[:, oy, xmap[cx][sel]] += 1
    xf = x.permute(0, 2, 3, 1).reshape(-1, Cin)                                             # pixel-major [B*Hin*Win, Cin]
    of = out.permute(0, 2, 3, 1).reshape(-1, Cout).clone()
    wf = writes.reshape(-1).clone()

    def list_pass(src, dst, n, sets, rowstep, colpitch_in, colpitch_out, c_in, c_nout, collive):
        c_up = c_nout + 4
        scale = np.float32(c_in) / np.float32(c_up)
        ci = np.minimum(np.floor(np.arange(c_up, dtype=np.float32) * scale).astype(np.int64), c_in - 1)
        for c in range(3):
            for i in range(n):
                s0, o0 = int(src[c, i]), int(dst[c, i])
                if s0 < 0:
                    assert o0 < 0
                    continue
                for oc in range(c_nout):
                    if collive is not None and not collive[oc]:
                        continue
                    acc = torch.zeros(Cout, dtype=torch.float64)
                    for d in range(3):
                        for k in range(5):
                            acc += sets[c][:, :, d, k] @ xf[s0 + d * rowstep + ci[oc + k] * colpitch_in]
                    of[o0 + oc * colpitch_out] = acc
                    wf[o0 + oc * colpitch_out] += 1
    list_pass(plan.row_src.numpy(), plan.row_out.numpy(), plan.row_n, rows, Win, 1, 1, Win, Wout, None)
    list_pass(plan.col_src.numpy(), plan.col_out.numpy(), plan.col_n, cols, 1, Win, Wout, Hin, Hout, plan.row_regular.numpy())
    assert torch.equal(wf, torch.ones_like(wf)), 'every output pixel must be written exactly once'
    got = of.reshape(B, Hout, Wout, Cout).permute(0, 3, 1, 2)
    assert torch.equal(got, ref)
    assert 9.0 <= plan.taps_per_output < 25.0


def test_analog_comparison_model_surface():
    """stereospike_b200.ann mirrors network/ANN_models.py + the ANN blocks: same state-dict keys and parameter count as the
    oracle restatement (which tests/test_oracle_vs_reference.py pins against the reference's own file), no CPU fallback."""
    import pytest
    import torch
    import stereospike_b200 as sb
    from oracle import ann_ref
    net, o = sb.ann.StereoSpike_equivalentANN(), ann_ref.AnalogUNet()
    assert sorted(net.state_dict()) == sorted(o.state_dict())
    assert net.count_trainable_params() == sum(p.numel() for p in o.parameters()) == 18158788
    assert isinstance(net.Ineurons, sb.neuron.IFNode) and net.Ineurons.v_threshold == float('inf')
    assert sb.ann.SteroSpike_equivalentANN is sb.ann.StereoSpike_equivalentANN
    net.increment_epoch()
    assert net.epoch == 1 and net.get_max_accuracy() == float('inf')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net(torch.zeros(1, 1, 4, 260, 346))


def _emulate_corr(plan, src, Cdst):
    """float64 emulation of ss_corr_bf16 as include/stereospike_b200.h specifies it: stride-1 correlation of the zero-padded source
    with every weight set on the virtual grid, result channel (tile, class, c) routed to destination channel tile * ntile + c at
    (ymap[row class][i], xmap[col class][j]); -1 = no destination; destinations accumulate."""
    g = plan.geom
    w = plan.w_img.double()                                   # (the test keeps the OIHW sets instead of the bf16 image)
    nt, nc = plan.ntile, plan.nclass
    r = F.conv2d(F.pad(src, (plan.pad,) * 4), w)              # [B, nsets * nt, Hv, Wv]
    B = src.shape[0]
    assert tuple(r.shape[2:]) == (plan.Hv, plan.Wv) and r.shape[1] == (Cdst // nt) * nc * nt
    r = r.view(B, Cdst // nt, nc, nt, plan.Hv, plan.Wv)
    dst = torch.zeros(B, Cdst, g.Hin, g.Win, dtype=torch.float64)
    if plan.ymap is None:
        assert nc == 1 and (plan.Hv, plan.Wv) == (g.Hin, g.Win)
        return r[:, :, 0].reshape(B, Cdst, g.Hin, g.Win)
    ym, xm = plan.ymap.view(-1, plan.Hv).long(), plan.xmap.view(-1, plan.Wv).long()
    for cls in range(nc):
        yy, xx = ym[cls // 2 if nc > 1 else 0], xm[cls % 2 if nc > 1 else 0]
        iy, ix = (yy >= 0).nonzero().flatten(), (xx >= 0).nonzero().flatten()
        part = r[:, :, cls].reshape(B, Cdst, plan.Hv, plan.Wv)[:, :, iy][:, :, :, ix]
        flat = (yy[iy][:, None] * g.Win + xx[ix][None, :]).reshape(-1)
        dst.view(B, Cdst, -1).index_add_(2, flat, part.reshape(B, Cdst, -1))
    return dst


@pytest.mark.parametrize('kind,Cin,Cout,ks,stride,pad,Hin,Win,up', [
    ('conv', 64, 32, 3, 1, 1, 17, 22, None),           # bottleneck convs
    ('conv', 32, 32, 5, 2, 2, 33, 44, None),           # conv4-type (odd height, even width)
    ('conv', 64, 16, 5, 2, 2, 26, 35, None),           # conv2-type sizes (even height, odd width), 64-channel weight sets
    ('upconv', 32, 16, 5, 1, 0, 17, 22, (33, 44)),     # deconv4
    ('upconv', 64, 32, 5, 1, 0, 13, 9, (27, 17)),
])
def test_data_gradient_plan_equals_autograd(monkeypatch, kind, Cin, Cout, ks, stride, pad, Hin, Win, up):
    """Host logic of ops.DgradPlan (correlation weights, parity classes of the stride-2 blocks, virtual grid and output maps of the
    upsampled blocks) run through a float64 emulation of the ss_corr_bf16 contract: equals autograd's input gradient of the
    reference block (Conv2d / UpsamplingNearest2d -> Conv2d, network/blocks.py:124-127)."""
    monkeypatch.setattr(ops, '_pack_bf16', lambda w_eff, k, ntile: w_eff.contiguous())      # keep the OIHW sets (no CUDA library call)
    gen = torch.Generator().manual_seed(Hin * 100 + Win)
    w = (torch.randn(Cout, Cin, ks, ks, generator=gen) / 8).double()                       # fp32-representable: the plan's .float() is lossless
    x = torch.randn(2, Cin, Hin, Win, generator=gen, dtype=torch.float64, requires_grad=True)
    if kind == 'conv':
        y = F.conv2d(x, w, stride=stride, padding=pad)
    else:
        y = F.conv2d(F.interpolate(x, size=(up[0] + ks - 1, up[1] + ks - 1), mode='nearest'), w)
        assert tuple(y.shape[2:]) == up
    gy = torch.randn(y.shape, generator=gen, dtype=torch.float64)
    want, = torch.autograd.grad(y, x, gy)
    geom = ops.BlockGeom(kind, Cin, Cout, ks, Hin, Win, int(y.shape[2]), int(y.shape[3]), stride, pad)
    plan = ops.DgradPlan(w, geom, 'cpu')
    assert plan.ntile == (64 if Cin % 64 == 0 else 32)
    got = _emulate_corr(plan, gy, Cin)
    assert float((got - want).abs().max()) < 1e-10


@pytest.mark.parametrize('Hin,Win,Hout,Wout', [(17, 22, 33, 44), (33, 44, 65, 87), (65, 87, 130, 173), (130, 173, 260, 346), (9, 11, 17, 21)])
def test_fold_plan_concatenated_lists_and_item_tables(Hin, Win, Hout, Wout):
    """The row-list passes read the CONCATENATED class lists (each class padded to whole 16-entry tiles on its own) through an item
    table {weight set, m-tile}: same entries per class as the padded lists the emulation above walks, every (weight set, tile) of a
    class exactly once, no tile of another class."""
    B = 3
    plan = ops.FoldPlan(Hin, Win, Hout, Wout, B, 'cpu')
    assert plan.ok
    for conc, src_pad, out_pad in ((plan.crow, plan.row_src, plan.row_out), (plan.ccol, plan.col_src, plan.col_out),
                                   (plan.creg, plan.reg_src, plan.reg_out)):
        src, out, n, ranges = conc
        assert n == src.numel() == out.numel() and n % 16 == 0
        assert len(ranges) == src_pad.shape[0]
        ty = 0
        for c, (ty0, nt) in enumerate(ranges):
            assert ty0 == ty
            ty += nt
            seg_s, seg_o = src[ty0 * 16:(ty0 + nt) * 16], out[ty0 * 16:(ty0 + nt) * 16]
            live = seg_s >= 0
            assert torch.equal(live, seg_o >= 0)
            assert bool(live[:int(live.sum())].all()), 'padding only at the end of a class'
            want = src_pad[c] >= 0
            assert torch.equal(seg_s[live], src_pad[c][want]) and torch.equal(seg_o[live], out_pad[c][want])
            assert nt == (int(live.sum()) + 15) // 16
        assert ty * 16 == n or (ty == 0 and n == 16)
    # every output row appears exactly once over the regular + irregular row lists, every column once over (dense map, column lists)
    rows = torch.cat([plan.crow[1][plan.crow[1] >= 0], plan.creg[1][plan.creg[1] >= 0]]).sort().values
    assert torch.equal(rows, (torch.arange(B * Hout) * Wout).int())
    cols0 = plan.ccol[1][plan.ccol[1] >= 0]
    dense_cols = plan.xmap[plan.xmap >= 0]
    assert torch.equal(torch.cat([cols0[cols0 < Wout], dense_cols]).sort().values, torch.arange(Wout).int())
    for which, col_classes in (('crow', 1), ('ccol', 1), ('creg', 2)):
        ranges = getattr(plan, which)[3]
        for ntiles_out, tiles_x in ((1, 1), (2, 3)):
            tab, n_items = plan.item_table(which, ntiles_out, tiles_x, col_classes)
            total_ty = sum(nt for _, nt in ranges)
            assert n_items == ntiles_out * col_classes * total_ty * tiles_x
            if n_items == 0:
                assert tab is None
                continue
            assert tuple(tab.shape) == (n_items, 2)
            seen = set()
            nclass = len(ranges) * col_classes
            for wset, mt in tab.tolist():
                assert (wset, mt) not in seen
                seen.add((wset, mt))
                rc = (wset % nclass) // col_classes
                ty0, nt = ranges[rc]
                assert 0 <= wset < ntiles_out * nclass and ty0 <= mt // tiles_x < ty0 + nt
