"""Shared GPU parity cases: the CUDA path against the oracle (oracle/sj_compat.py, oracle/ref_model.py) on the same
seeded inputs.  Used by tests/test_gpu_parity.py and tools/gpu_bringup.py."""
import time


def block_case(kind, Cin, Cout, ks, Hin, Win, stride, pad, up, neuron, T, B, impl, planes, resid, gain=3.0, seed=0):
    import torch
    import torch.nn.functional as F
    from oracle import sj_compat as sj
    from stereospike_b200 import ops
    from stereospike_b200._lib import SS_IN_U8_TBHWC
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(seed)
    if kind == 'conv':
        Hout, Wout = ops.conv_out_size(Hin, ks, stride, pad), ops.conv_out_size(Win, ks, stride, pad)
        geom = ops.BlockGeom('conv', Cin, Cout, ks, Hin, Win, Hout, Wout, stride, pad)
    else:
        Hout, Wout = up
        geom = ops.BlockGeom('upconv', Cin, Cout, ks, Hin, Win, Hout, Wout)
    x = (torch.rand(T, B, Cin, Hin, Win, generator=g) < 0.15).float() * torch.randint(1, 3, (T, B, Cin, Hin, Win), generator=g).float()
    w = (torch.rand(Cout, Cin, ks, ks, generator=g) * 2 - 1) / (Cin * ks * ks) ** 0.5
    r = (torch.rand(T, B, Cout, Hout, Wout, generator=g) < 0.3).float() if resid else None
    # oracle (CPU fp32)
    if neuron == 0:
        node = sj.IFNode(1.0, 0.0, sj.ATan(), True)
    elif neuron == 1:
        node = sj.LIFNode(3.0, 1.0, 0.0, sj.ATan(), True)
    else:
        node = sj.ParametricLIFNode(3.0, 1.0, 0.0, sj.Sigmoid(), True)
    hs, outs = [], []
    with torch.no_grad():
        for t in range(T):
            xi = x[t]
            if kind == 'upconv':
                xi = F.interpolate(xi, size=(Hout + ks - 1, Wout + ks - 1), mode='nearest')
                y = F.conv2d(xi, w)
            else:
                y = F.conv2d(xi, w, stride=stride, padding=pad)
            y = y * gain
            node.neuronal_charge(y)
            hs.append(node.v.clone())
            node.neuronal_fire()
            node.neuronal_reset()
            o = node.spike.clone()
            if resid:
                o = o + r[t]
            outs.append(o)
    h_ref = torch.stack(hs).permute(0, 1, 3, 4, 2).contiguous()
    o_ref = torch.stack(outs).permute(0, 1, 3, 4, 2).contiguous()
    v_ref = node.v.permute(0, 2, 3, 1).contiguous()
    # device
    xb = x.permute(0, 1, 3, 4, 2).contiguous().to(dev, torch.uint8)
    cin_dev = None
    if Cin < 4:       # first-layer mode wants 4 packed channels
        xb = torch.cat([xb, xb.new_zeros(T, B, Hin, Win, 4 - Cin)], dim=-1).contiguous()
        cin_dev = 4
    rb = r.permute(0, 1, 3, 4, 2).contiguous().to(dev, torch.uint8) if resid else None
    decay = node.w.detach().sigmoid().reshape(1).float().to(dev) if neuron == 2 else None
    common = dict(T=T, B=B, neuron=neuron, gain=gain, v_th=1.0, v_reset=0.0, tau=3.0, decay=decay, want_v_out=True,
                  resid=rb, want_h=True)
    if impl == 'umma':
        w_i8, wscale, _ = ops.pack_weights_i8(w.to(dev), planes, cin_pad=cin_dev)
        torch.cuda.synchronize()
        t0 = time.time()
        out, v_out, h_seq = ops.conv_i8_fwd(xb, geom, w_i8, wscale, planes=planes, cin=cin_dev, **common)
    else:
        w_kn = ops.weight_to_kn(w.to(dev))
        torch.cuda.synchronize()
        t0 = time.time()
        out, v_out, h_seq = ops.conv_neuron_fwd(xb, geom, w_kn, in_layout=SS_IN_U8_TBHWC, **common)
    torch.cuda.synchronize()
    dt = time.time() - t0
    h = h_seq.cpu()
    o = out.float().cpu()
    dh = (h - h_ref).abs()
    band = (h_ref - 1.0).abs() > 1e-4
    mism = ((o != o_ref) & band).sum().item()
    mism_all = (o != o_ref).sum().item()
    res = dict(max_dh=float(dh.max()), mean_dh=float(dh.mean()), h_absmax=float(h_ref.abs().max()),
               spike_mismatch_outside_band=mism, spike_mismatch_all=mism_all, n=o.numel(),
               rate=float((o_ref > 0).float().mean()), v_maxdiff=float((v_out.cpu() - v_ref).abs().max()), ms=dt * 1e3)
    # first-timestep error alone (teacher-forcing proxy: later steps inherit state differences after a flip)
    res['max_dh_t0'] = float(dh[0].max())
    return res


def model_case(variant, mono, gain, T, B, impl, planes, seed=11, backward=False, tau=3.0, with_fp64=False, bwd_impl=None,
               fold_min_frames=None):
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    from oracle.make_golden import simple_loss
    import stereospike_b200 as sb
    torch.manual_seed(seed)
    o = rm.SpikingUNet(variant, mono, surrogate_function=sj.ATan() if variant == 'if' else None, tau=tau, multiply_factor=gain)
    if variant == 'if':
        n = sb.StereoSpike(surrogate_function=sb.surrogate.ATan(), multiply_factor=gain)
    elif mono:
        n = sb.fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike(use_plif=variant == 'plif', tau=tau, multiply_factor=gain)
    else:
        n = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=variant == 'plif', tau=tau, multiply_factor=gain)
    n.load_state_dict(o.state_dict())
    n = n.cuda()
    n.set_kernel_options(impl=impl, weight_planes=planes, bwd_impl=bwd_impl, fold_min_frames=fold_min_frames)
    x = rm.synthetic_inputs(B, T, 2 if mono else 4, seed=seed + 1)
    label = rm.synthetic_label(B, seed=seed + 2)
    sj.reset_net(o)
    sb.functional.reset_net(n)
    ctx = torch.enable_grad() if backward else torch.no_grad()
    with ctx:
        t0 = time.time()
        ref = o.forward_seq(x, return_all=True)
        t_cpu = time.time() - t0
        d_ref, layers_ref = ref[0], ref[2]
        xg = x.cuda()
        torch.cuda.synchronize()
        t0 = time.time()
        got = n.forward_seq(xg)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        d_got = got if mono else got[0]
        res = {'t_cpu_s': t_cpu, 't_gpu_first_s': t_gpu}
        for i, (a, b) in enumerate(zip(d_ref, d_got)):
            diff = (a.detach() - b.detach().cpu()).abs()
            res[f'depth{i + 1}_max'] = float(diff.max())
            res[f'depth{i + 1}_mean'] = float(diff.mean())
        res['depth_absmean'] = float(d_ref[0].detach().abs().mean())
        res['mde_ref'] = float(rm.mean_depth_error(d_ref[0].detach(), label))
        if with_fp64:
            # the same oracle evaluated in float64: |mde_ref - mde_ref64| measures how chaotic this configuration is
            import copy
            o64 = copy.deepcopy(o).double()
            sj.reset_net(o64)
            with torch.no_grad():
                r64 = o64.forward_seq(x.double(), return_all=True)
            res['mde_ref64'] = float(rm.mean_depth_error(r64[0][0].float(), label))
        res['mde_got'] = float(rm.mean_depth_error(d_got[0].detach().cpu(), label))
        acts = n.engine  # layer mismatch rates from the engine's last activations
        side_acts = None
        if backward:
            simple_loss(d_ref, label).backward()
            simple_loss(d_got, label.cuda()).backward()
            go = dict(o.named_parameters())
            worst = (1.0, None)
            rel = {}
            for k, p in n.named_parameters():
                a, b = go[k].grad.flatten().double(), p.grad.detach().cpu().flatten().double()
                cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
                rel[k] = (cos, float((a - b).norm() / (a.norm() + 1e-30)))
                if cos < worst[0]:
                    worst = (cos, k)
            res['grad_worst_cos'] = worst
            res['grad_rel'] = {k: (round(v[0], 6), round(v[1], 5)) for k, v in rel.items()}
    # per-layer spike mismatch (needs a second no-grad run to fetch activations)
    with torch.no_grad():
        sb.functional.reset_net(n)
        _, side = n.engine.run(x.cuda())
        mm = {}
        for k, v in layers_ref.items():
            if k.startswith('out_deconv'):
                continue
            a = side['acts'][k][-1].permute(0, 3, 1, 2).float().cpu()
            mm[k] = (float((a != v.detach()).float().mean()), float((v.detach() != 0).float().mean()))
        res['mismatch(rate,firing)'] = {k: (round(a, 7), round(b, 4)) for k, (a, b) in mm.items()}
    return res




# ------------------------------------------------------------------------------------------------------------------
# Teacher-forced parity of EVERY block of a model at its real geometry (the benchmark configuration included): chaos
# cannot touch this test -- each CUDA block gets the oracle's own input spikes, so a threshold flip upstream never
# reaches it (SURVEY.md section 8(c) protocol (i)).
def _site_nodes(o):
    """engine site name -> the oracle's neuron module of that block."""
    nodes = {'bottom': o.bottom[2]}
    for k in ('conv1', 'conv2', 'conv3', 'conv4', 'deconv4', 'deconv3', 'deconv2', 'deconv1'):
        nodes[k] = getattr(o, k)[2]
    for i, blk in enumerate(o.bottleneck):
        nodes[f'bottleneck.{i}.conv1'] = blk.sn1
        nodes[f'bottleneck.{i}.conv2'] = blk.sn2
    return nodes


def oracle_trace(o, x):
    """Runs the oracle timestep by timestep (state must be reset by the caller).  Returns (depths of the last call,
    acts: engine activation name -> u8 [T,B,H,W,C], h: engine site name -> fp32 [T,B,H,W,C] pre-reset potentials)."""
    import torch
    T = int(x.shape[1])
    nodes = _site_nodes(o)
    h = {k: [] for k in nodes}
    undo = []
    for name, node in nodes.items():
        orig = node.neuronal_fire

        def fire(node=node, orig=orig, store=h[name]):
            store.append(node.v.detach().permute(0, 2, 3, 1).contiguous())
            orig()
        node.neuronal_fire = fire
        undo.append(node)
    mids = {}
    hooks = []
    for i, blk in enumerate(o.bottleneck):
        # sn1's output feeds conv2 (captured before anything can touch it); the block output is `sn2(...) += x`
        hooks.append(blk.sn1.register_forward_hook(
            lambda m, a, out, k=f'_sew{i}_mid': mids.setdefault(k, []).append(out.detach().clone())))
        if i + 1 < len(o.bottleneck):
            hooks.append(blk.register_forward_hook(
                lambda m, a, out, k=f'_sew{i}_out': mids.setdefault(k, []).append(out.detach().clone())))
    acts = {}
    try:
        with torch.no_grad():
            for t in range(T):
                depths, _, layers = o.forward(x[:, t:t + 1], return_all=True)
                for k, v in layers.items():
                    acts.setdefault(k, []).append(v.detach().clone())
    finally:
        for node in undo:
            del node.neuronal_fire
        for hk in hooks:
            hk.remove()
    acts.update(mids)
    to_u8 = lambda lst: torch.stack(lst).permute(0, 1, 3, 4, 2).contiguous().to(torch.uint8)
    return depths, {k: to_u8(v) for k, v in acts.items()}, {k: torch.stack(v) for k, v in h.items()}


def build_pair(variant, mono, gain, tau, seed, planes=3, impl='umma', in_channels=None):
    """(oracle, CUDA model) with identical weights."""
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    import stereospike_b200 as sb
    torch.manual_seed(seed)
    o = rm.SpikingUNet(variant, mono, surrogate_function=sj.ATan() if variant == 'if' else None, tau=tau, multiply_factor=gain,
                       in_channels=in_channels)
    kw = {} if in_channels is None else {'in_channels': in_channels}
    if variant == 'if':
        n = sb.StereoSpike(surrogate_function=sb.surrogate.ATan(), multiply_factor=gain, **kw)
    elif mono:
        n = sb.fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike(use_plif=variant == 'plif', tau=tau, multiply_factor=gain, **kw)
    else:
        n = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=variant == 'plif', tau=tau, multiply_factor=gain, **kw)
    n.load_state_dict(o.state_dict())
    n = n.cuda()
    n.set_kernel_options(impl=impl, weight_planes=planes)
    return o, n


def teacher_forced_model(o, n, x, trace=None, planes=3, band=1e-4, fold=True):
    """Every fused block of CUDA model ``n`` fed with the ORACLE's input spikes of that block (all T timesteps, real
    geometry) and compared with the oracle's pre-reset potentials.  A neuron is compared at step t only while its spikes
    agreed at every earlier step (after a flip inside the threshold band its state legitimately differs).
    Returns {site: dict(max_dh, h_absmax, flips, flips_outside_band, out_mismatch_outside_band, n, rate)}."""
    import torch
    from oracle import sj_compat as sj
    from stereospike_b200 import ops
    dev = torch.device('cuda')
    if trace is None:
        sj.reset_net(o)
        trace = oracle_trace(o, x)
    _, acts, hs = trace
    B, T = int(x.shape[0]), int(x.shape[1])
    eng = n.engine
    res = {}
    for s in eng.sites:
        first = s.src == 'x'
        node = s.node
        if first:
            xin = ops.pack_events(x.to(dev))
            Hin, Win = int(x.shape[3]), int(x.shape[4])
        else:
            xin = acts[s.src].to(dev)
            Hin, Win = int(xin.shape[2]), int(xin.shape[3])
        g = s.geom(Hin, Win)
        resid = acts[s.resid].to(dev) if s.resid is not None else None
        use_fold = bool(fold) and g.kind == 'upconv' and g.ks == 5
        _, w_i8 = s.packed(planes, True, need_kn=False, fold=use_fold)
        decay = node.decay_tensor()
        kw = dict(T=T, B=B, neuron=node.kind, gain=s.gain_mod.gain(), v_th=node.v_threshold, v_reset=node.v_reset,
                  tau=node._tau_value(), decay=decay.detach().contiguous() if decay is not None else None, resid=resid,
                  want_h=True, planes=planes)
        if use_fold:
            out, _, h_got = ops.conv_i8_fwd_folded(xin, g, w_i8[0], w_i8[1], w_i8[2], w_i8[3], **kw)
        else:
            out, _, h_got = ops.conv_i8_fwd(xin, g, w_i8[0], w_i8[1], cin=ops.first_layer_channels(g.Cin) if first else g.Cin, **kw)
        h_ref = hs[s.name].to(dev)
        vth = float(node.v_threshold)
        s_ref, s_got = h_ref >= vth, h_got >= vth
        valid = torch.ones_like(s_ref[0])
        max_dh, flips, flips_out = 0.0, 0, 0
        for t in range(T):
            d = ((h_got[t] - h_ref[t]).abs() * valid).max()
            max_dh = max(max_dh, float(d))
            flip = (s_ref[t] != s_got[t]) & valid
            flips += int(flip.sum())
            flips_out += int((flip & ((h_ref[t] - vth).abs() > band)).sum())
            valid = valid & ~flip
        # the block's stored output (spikes + residual) against the oracle's, for neurons that never flipped
        o_ref = acts[s.out].to(dev)
        bad = ((out != o_ref) & valid.unsqueeze(0)).sum()
        res[s.name] = dict(max_dh=max_dh, h_absmax=float(h_ref.abs().max()), flips=flips, flips_outside_band=flips_out,
                           out_mismatch_outside_band=int(bad), n=int(h_ref.numel()), rate=float(s_ref.float().mean()))
        del h_ref, h_got, out, o_ref, xin, resid
    return res


def parity_summary(variant='lif', gain=15.0, tau=3.0, T=5, B=1, seed=0, planes=3, x_seed=100, with_fp64=True, fold=True):
    """End-to-end + teacher-forced parity of one configuration as a small dict (bench.py prints it; the GPU tests assert on
    it).  ``oracle_fp32_vs_fp64`` is the oracle's own sensitivity: |MDE(fp32) - MDE(float64)| of the same reference code."""
    import copy
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    import stereospike_b200 as sb
    o, n = build_pair(variant, False, gain, tau, seed, planes)
    n.set_kernel_options(fold_upsample=bool(fold), fold_min_frames=1)      # the slice runs the kernels of the benchmarked batch
    x = rm.synthetic_inputs(B, T, 4, seed=x_seed)
    label = rm.synthetic_label(B, seed=x_seed + 1)
    sj.reset_net(o)
    trace = oracle_trace(o, x)
    d_ref = trace[0]
    with torch.no_grad():
        sb.functional.reset_net(n)
        depths, side = n.engine.run(x.cuda())
    d1 = depths[3].unsqueeze(1).cpu()
    mde_ref = float(rm.mean_depth_error(d_ref[0], label))
    mde_got = float(rm.mean_depth_error(d1, label))
    out = {'config': f'{variant} gain {gain} tau {tau} T={T} B={B} planes={planes}', 'mde_oracle_fp32': mde_ref, 'mde_cuda': mde_got,
           'mde_abs_diff': abs(mde_ref - mde_got), 'depth_mean_abs_diff': float((d_ref[0] - d1).abs().mean())}
    worst = (0.0, None)
    for k, a_ref in trace[1].items():
        if k.startswith('out_deconv') or k not in side['acts']:
            continue
        mm = float((side['acts'][k].cpu() != a_ref).float().mean())
        if mm >= worst[0]:
            worst = (mm, k)
    out['worst_layer_mismatch'] = {'layer': worst[1], 'rate': worst[0]}
    if with_fp64:
        o64 = copy.deepcopy(o).double()
        sj.reset_net(o64)
        with torch.no_grad():
            d64 = o64.forward_seq(x.double())[0]
        mde64 = float(rm.mean_depth_error(d64[0].float(), label))
        out['mde_oracle_fp64'] = mde64
        out['oracle_fp32_vs_fp64'] = abs(mde_ref - mde64)
        out['mde_abs_diff_vs_fp64'] = abs(mde64 - mde_got)
    tf = teacher_forced_model(o, n, x, trace=trace, planes=planes, fold=fold)
    out['teacher_forced'] = {'max_dh': max(v['max_dh'] for v in tf.values()),
                             'flips_outside_band': sum(v['flips_outside_band'] for v in tf.values()),
                             'flips_in_band': sum(v['flips'] for v in tf.values()),
                             'neuron_steps': sum(v['n'] for v in tf.values())}
    return out
