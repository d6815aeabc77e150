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


def model_case(variant, mono, gain, T, B, impl, planes, seed=11, backward=False, tau=3.0, with_fp64=False, bwd_impl=None):
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
    n.set_kernel_options(impl=impl, weight_planes=planes, bwd_impl=bwd_impl)
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


