#!/usr/bin/env python
"""Headline benchmark: event-frames/sec of the fused spiking U-Net forward (binocular, T=5, 260x346, batch 8 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of the whole path over one batch of synthetic event frames: B*T event-frames.
Rank 0 prints ONE JSON line (contract in the task statement / DESIGN.md section "Measurement").
  value      device-timed throughput, inputs resident in HBM (rotating input sets larger than L2)
  e2e        same metric through the public API (pipeline.HostPipeline.step) with pinned-HOST inputs: H2D of the step's
             frames and D2H of the finest depth map inside the timed region
  roofline   tcgen05 conv+neuron kernel: algorithmic FLOPs / CUDA-event time of those launches vs the measured BURST
             bf16 peak (MEASURED_PEAKS.json; the sustained figure is printed beside it)
  parity     (N=1) the CUDA path against the oracle on a B=1 slice of the same configuration, outside the timed region:
             |dMDE|, the oracle's own fp32-vs-float64 sensitivity, worst per-layer spike mismatch, teacher-forced max |dh|
  train      the training step of BASELINE.json configs[2]/[3] (B=16 per GPU, T=5: forward + surrogate backward + NCCL
             gradient all-reduce overlapped with the backward + Adam), a short run after the inference measurement
  cpu_baseline  the oracle (pure-PyTorch restatement of the reference, oracle/ref_model.py) on the host cores,
             bounded sample
`--impl reference` times that CPU oracle alone (the reference cannot be pip-installed: it has no packaging and its
neuron library is absent -- DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# The training step allocates ~9 GB of transient tensors per step on two streams; with the default caching allocator a fragmented
# pool occasionally grows by a cudaMalloc -- a device synchronisation that costs 3-17 ms on that step (tools/train_step_trace.py:
# 1-2 such steps in 40).  Expandable segments grow the pool without it.  Must be set before torch initialises CUDA.
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H0, W0 = 260, 346
# forward conv FLOPs per event-frame of every tensor-core block (2*M*N*K at reference geometry)
MFLOP_PER_FRAME = {'bottom': 575.7, 'conv1': 2303.0, 'conv2': 2316.3, 'conv3': 2379.0, 'conv4': 2451.0,
                   'bottleneck.0.conv1': 1764.8, 'bottleneck.0.conv2': 1764.8, 'bottleneck.1.conv1': 1764.8,
                   'bottleneck.1.conv2': 1764.8, 'deconv4': 9515.8, 'deconv3': 9265.2, 'deconv2': 9211.9,
                   'deconv1': 9211.9, 'heads': 777.2}
TOTAL_GFLOP_PER_FRAME = sum(MFLOP_PER_FRAME.values()) / 1e3
FOLD_DEFAULT = 1      # decoder blocks folded to 3x3 convs on the source (engine default)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return {'bf16_sustained': d.get('bf16_tflops_sustained', 1400.0), 'bf16_burst': d.get('bf16_tflops', 1590.0),
                'hbm': d.get('hbm_gbs', 6650.0), 'source': 'measured'}
    return {'bf16_sustained': 1400.0, 'bf16_burst': 1590.0, 'hbm': 6650.0, 'source': 'fallback'}


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '20'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Only samples taken after this call count (the timed region)."""
        self.f.flush()
        try:
            self.skip = sum(1 for _ in open(self.f.name))
        except OSError:
            self.skip = 0

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        allrows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        skip = getattr(self, 'skip', 0)
        rows, window = allrows[skip:], 'timed region'
        if not rows and allrows:
            # a timed region shorter than one 20 ms poll: fall back to the last samples of the warm-up (same kernels, same load)
            rows, window = allrows[max(0, skip - 3):], 'end of warm-up (timed region shorter than one poll)'
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower() == 'active':
                    reasons.add(n)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def build_oracle(neuron, gain, tau):
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    torch.manual_seed(0)
    if neuron == 'if':
        return rm.SpikingUNet('if', surrogate_function=sj.ATan(), multiply_factor=gain)
    return rm.SpikingUNet(neuron, tau=tau, multiply_factor=gain)


def cpu_oracle_rate(neuron, gain, tau, T, sample_B, reps):
    """event-frames/s of the CPU oracle on all host threads (forward only, no_grad)."""
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = build_oracle(neuron, gain, tau)
    x = rm.synthetic_inputs(sample_B, T, 4, seed=0)
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            sj.reset_net(net)
            t0 = time.perf_counter()
            net.forward_seq(x)
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
    best = statistics.median(times)
    return sample_B * T / best, cores, best


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = build_oracle(args.neuron, args.gain, args.tau)
    on_gpu = args.reference_device == 'cuda'
    sample_B = args.batch                  # the batch the config declares (the arm ran B=1 while printing batch 8)
    x = rm.synthetic_inputs(sample_B, args.T, 4, seed=0)
    label = rm.synthetic_label(sample_B, seed=1)
    if on_gpu:
        # what the reference itself executes on cuda:0: cuDNN convs + unfused elementwise neuron kernels (true fp32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        net, x, label = net.cuda(), x.cuda(), label.cuda()
    opt = torch.optim.Adam(net.parameters(), lr=2e-4) if args.mode == 'train' else None

    def ref_step():
        sj.reset_net(net)
        if args.mode == 'train':
            depths = net.forward_seq(x)[0]
            masked_l1(depths, label).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        else:
            with torch.no_grad():
                net.forward_seq(x)

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        ref_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_step()
    sync()
    dt = time.perf_counter() - t0
    val = sample_B * args.T * args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'event-frames/sec', 'value': val, 'unit': 'event-frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, per_gpu_B=args.batch),
        'cpu_baseline': {'value': val, 'unit': 'event-frames/s', 'cores': cores, 'kind': 'port',
                         'sample': f'each step = the full B={sample_B} x T={args.T} batch of the same workload, {args.mode}, '
                                   f'torch {torch.__version__} ' + ('CUDA eager (cuDNN fp32, allow_tf32=False) -- the reference\'s own GPU '
                                   'execution model, NOT the CPU arm' if on_gpu else f'CPU fp32, {cores} threads') +
                                   '; one process whatever --gpus says (the reference has no distributed code)'},
        'reference_device': args.reference_device,
        'e2e': {'value': val, 'unit': 'event-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def masked_l1(depths, label):
    """Sum over the four scales of the NaN-masked mean absolute error (network/metrics.py:83-95 per scale); stands in for
    network/loss.py::Total_Loss, which cannot run on CPU tensors on a GPU box."""
    import torch
    mask = ~torch.isnan(label)
    n = mask.count_nonzero()
    lab = torch.nan_to_num(label)
    tot = 0.0
    for d in depths:
        tot = tot + ((d - lab).abs() * mask).sum() / n
    return tot


def workload_config(args, per_gpu_B):
    return {'workload': f'StereoSpike spiking U-Net forward (fused conv+{args.neuron.upper()} blocks + heads/I-neurons), '
                        f'binocular 4x{H0}x{W0} event frames, T={args.T}, batch {per_gpu_B} per GPU, fp32-parity ' + ('inference ' if args.mode == 'infer' else 'training step (forward + surrogate backward + Adam) ') +
                        f'(u8 spikes x {args.planes} int8 weight digit planes on the int8 tensor cores, exact s32 accumulate, one fp32 rounding)',
            'mode': args.mode, 'neuron': args.neuron, 'T': args.T, 'batch_per_gpu': per_gpu_B, 'global_batch': per_gpu_B * args.gpus,
            'weight_planes': args.planes, 'multiply_factor': args.gain, 'tau': args.tau,
            'state': 'reset before every step (test.py per-sample protocol); final membrane potentials ' +
                     ('kept (keep_state)' if args.keep_state else 'not written back (stateless serving, keep_state=False)'),
            'fold_upsample': bool(args.fold),
            'l2': f'rotating {args.input_sets} input sets per step; per-step activation stream (~2 GB) exceeds the 126 MB L2',
            'parallelism': f'replicas x{args.gpus} (batch shards, no data-path collective)' if args.mode == 'infer' else
                           f'dp{args.gpus} (batch shards, NCCL gradient all-reduce issued per block inside the backward)'}


def make_net(args, dev):
    import torch
    import stereospike_b200 as sb
    torch.manual_seed(0)
    if args.neuron == 'if':
        net = sb.StereoSpike(surrogate_function=sb.surrogate.ATan(), multiply_factor=args.gain)
    else:
        net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=args.neuron == 'plif', tau=args.tau,
                                                                              multiply_factor=args.gain)
    net = net.to(dev)
    net.set_kernel_options(impl=args.kernel, weight_planes=args.planes, fold_upsample=bool(args.fold))
    return net


def time_training(args, dev, world, rank, B, steps, warmup, barrier):
    """One training configuration (BASELINE.json configs[2] at N=1, configs[3] at N=8): forward + fused loss + surrogate
    backward with the per-block gradient all-reduce overlapped (parallel.OverlappedGradientSync) + Adam.  Returns
    (ms per step with the all-reduce, ms per step with the all-reduce switched off, collectives per step, launches per step)."""
    import torch
    import stereospike_b200 as sb
    from stereospike_b200 import _lib
    from oracle import ref_model as rm          # synthetic-input recipe only
    net = make_net(args, dev)
    T = args.T
    xs = [rm.synthetic_inputs(B, T, 4, seed=500 + rank * 16 + i).to(dev) for i in range(2)]
    label = rm.synthetic_label(B, seed=1).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=2e-4)
    crit = sb.loss.Total_Loss()
    sync = sb.parallel.OverlappedGradientSync(net)
    sync.set_batch(local_samples=B, global_samples=B * world)

    def step(x):
        sb.functional.reset_net(net)
        out = net.forward_seq(x)
        crit(out[0], label).backward()          # the all-reduces are issued from inside this backward, block by block
        opt.step()
        opt.zero_grad(set_to_none=True)

    res = []
    for use_sync in (True, False):
        if use_sync and world > 1:
            sync.attach()
        else:
            sync.detach()
        for i in range(warmup):
            step(xs[i % 2])
        barrier()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(xs[i % 2])
        e1.record()
        barrier()
        res.append((e0.elapsed_time(e1) / steps, (_lib.launch_count() - l0) / steps))
        if world == 1:
            res.append(res[0])
            break
    sync.detach()
    ncoll = sync.collectives
    del net, opt, xs
    torch.cuda.empty_cache()
    return res[0][0], res[1][0], ncoll, res[0][1]


def sj_cupy_proxy(args, dev, B, T, steps=10, warmup=3):
    """PROXY for the reference model on SpikingJelly's multi-step cupy backend (SURVEY.md 8(d) comparison point 2; neither
    spikingjelly nor cupy exists on the box, so this is NOT that library): what that path executes on a GPU -- cuDNN convolutions
    over the flattened [T*B] batch (fp32 NCHW tensors), an elementwise gain, and ONE unfused per-neuron T-loop CUDA kernel per layer
    (one thread per neuron, looping over T, reading the fp32 conv output and writing fp32 spikes: ss_neuron_fwd, the same structure
    as MultiStepLIFNode's cupy kernel), elementwise skip adds, upsample + conv heads.  Forward only.  Returns event-frames/s for
    cuDNN in true fp32 and with TF32 allowed (PyTorch's default, which is what an unmodified SpikingJelly script gets)."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from stereospike_b200 import _lib
    from oracle import ref_model as rm
    L = _lib.lib()
    net = build_oracle(args.neuron, args.gain, args.tau).to(dev)
    x = rm.synthetic_inputs(B, T, 4, seed=900).to(dev)
    kind = {'if': 0, 'lif': 1, 'plif': 2}[args.neuron]
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def multistep(node, y):          # y fp32 [T*B, C, H, W] -> spikes, one T-loop kernel
        yt = y.view(T, -1)
        s = torch.empty_like(yt)
        v = torch.zeros(yt.shape[1], dtype=torch.float32, device=dev)
        k, decay, tau = kind, None, float(getattr(node, 'tau', 2.0))
        if hasattr(node, 'w'):
            k, decay = 2, node.w.detach().sigmoid().reshape(1).float()
        elif not hasattr(node, 'tau'):
            k = 0
        _lib.check(L.ss_neuron_fwd(T, yt.shape[1], k, 1.0, 0.0, tau, vp(decay), vp(yt), vp(v), vp(s), None, stream()), 'ss_neuron_fwd')
        return s.view_as(y)

    def block(seq, xin):             # Sequential(conv | UpConv, Gain, neuron)
        return multistep(seq[2], seq[1](seq[0](xin)))

    def fwd():
        f = x.transpose(0, 1).reshape(T * B, 4, H0, W0)          # time-major flattened batch
        b = block(net.bottom, f)
        c1 = block(net.conv1, b); c2 = block(net.conv2, c1); c3 = block(net.conv3, c2); c4 = block(net.conv4, c3)
        r = c4
        for blk in net.bottleneck:
            m = multistep(blk.sn1, blk.conv1(r))
            r = multistep(blk.sn2, blk.conv2(m)) + r
        cur, depth = r, 0.0
        for (name, _, _, _), skip, (hname, _) in zip(rm.DEC, (c3, c2, c1, b), rm.HEADS):
            cur = block(getattr(net, name), cur) + skip
            depth = depth + getattr(net, hname)(cur).view(T, B, 1, H0, W0).sum(0)       # I-neurons: never fire, potentials add up
        return depth

    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            for _ in range(warmup):
                fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fwd()
            e1.record()
            torch.cuda.synchronize()
        out['tf32' if tf32 else 'fp32'] = B * T * steps / (e0.elapsed_time(e1) / 1e3)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    del net, x
    torch.cuda.empty_cache()
    return out


def timestep_sweep(args, dev, B=16, Ts=(1, 5, 10, 20), steps=10, warmup=3):
    """BASELINE.json configs[4]: T in {1, 5, 10, 20} at batch 16 on one GPU, stateless inference, device-resident frames."""
    import torch
    import stereospike_b200 as sb
    from oracle import ref_model as rm
    net = make_net(args, dev)
    net.set_kernel_options(keep_state=False)
    res = {}
    for T in Ts:
        xs = [rm.synthetic_inputs(B, T, 4, seed=700 + i).to(dev) for i in range(2)]
        with torch.no_grad():
            for i in range(warmup):
                sb.functional.reset_net(net)
                net.forward_seq(xs[i % 2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                sb.functional.reset_net(net)
                net.forward_seq(xs[i % 2])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[f'T={T}'] = {'ms_per_step': ms, 'event_frames_per_s': B * T / (ms / 1e3), 'depth_maps_per_s': B / (ms / 1e3)}
        del xs
        torch.cuda.empty_cache()
    del net
    torch.cuda.empty_cache()
    return {'batch': B, 'steps': steps, 'warmup': warmup, 'results': res}


def analog_model_record(dev, B=8, steps=5, warmup=3):
    """The analog comparison model (network/ANN_models.py, one frame per depth map): eval-mode forward on the library's fp32
    convolution kernels, beside the same model in PyTorch eager on the same GPU (cuDNN: true fp32 and TF32 allowed) and the CPU
    oracle.  It is the paper's Table-4 baseline, not the spiking hot path: the convolutions run on CUDA cores in exact fp32."""
    import time
    import torch
    import stereospike_b200 as sb
    from oracle import ann_ref, ref_model as rm
    torch.manual_seed(5)
    oracle = ann_ref.AnalogUNet().eval()
    ann_ref.randomize_batchnorm(oracle, seed=6)
    net = sb.ann.StereoSpike_equivalentANN()
    net.load_state_dict(oracle.state_dict())
    net = net.to(dev).eval()
    x = rm.synthetic_inputs(B, 1, 4, seed=31)
    xd = x.to(dev)

    def timed(fn):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def ours():
        sb.functional.reset_net(net)
        return net(xd)

    ms = timed(ours)
    with torch.no_grad():
        t0 = time.perf_counter()
        want = oracle(x[:1])
        cpu_s = time.perf_counter() - t0
        oracle.reset()
        sb.functional.reset_net(net)
        got = net(xd[:1])
    diff = max(float((g.cpu() - w).abs().max()) for g, w in zip(got, want)) / max(float(w.abs().max()) for w in want)
    eager = oracle.to(dev)
    rec = {'value': B / (ms / 1e3), 'unit': 'frames/s', 'ms_per_step': ms, 'batch': B, 'mode': 'eval forward',
           'max_depth_diff_vs_oracle_rel': diff, 'cpu_oracle_frames_per_s': 1.0 / cpu_s}

    def eager_fn():
        eager.reset()
        return eager(xd)

    before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    for name, tf32 in (('torch_eager_fp32_frames_per_s', False), ('torch_eager_tf32_frames_per_s', True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        rec[name] = B / (timed(eager_fn) / 1e3)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = before
    del net, eager, oracle
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import stereospike_b200 as sb
    from stereospike_b200 import _lib
    from stereospike_b200.pipeline import pack_events_host
    from oracle import ref_model as rm          # synthetic-input recipe only (shared with the CPU arm)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the stereospike_b200 hot path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.lib()   # fail loudly if the extension is missing
    sampler = ClockSampler(local) if rank == 0 else None      # started early so that it has samples under load

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    train = args.mode == 'train'
    net = make_net(args, dev)
    net.set_kernel_options(keep_state=bool(args.keep_state) or train)
    B, T = args.batch, args.T
    xs_f32 = [rm.synthetic_inputs(B, T, 4, seed=100 + rank * 16 + i) for i in range(args.input_sets)]
    # host frames as the data loader hands them over: packed u8 counts [T,B,H,W,4] (14.4 MB per batch) by default,
    # the reference's fp32 [B,T,4,H,W] tensors (57.6 MB) with --host-frames f32
    if args.host_frames == 'u8' and not train:
        xs_host = [pack_events_host(x).pin_memory() for x in xs_f32]
    else:
        xs_host = [x.pin_memory() for x in xs_f32]
    xs = [x.to(dev) for x in xs_f32]            # device-resident arm: the reference's fp32 frames, packed by ss_pack_events each step
    del xs_f32

    if train:
        label = rm.synthetic_label(B, seed=1).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=2e-4)
        crit = sb.loss.Total_Loss()
        gsync = sb.parallel.OverlappedGradientSync(net)
        gsync.set_batch(local_samples=B, global_samples=B * world)
        if world > 1:
            gsync.attach()

    def step(x):
        sb.functional.reset_net(net)
        if train:
            out = net.forward_seq(x)
            crit(out[0], label).backward()      # NCCL all-reduce of each block's gradient is issued inside the backward
            opt.step()
            opt.zero_grad(set_to_none=True)
            return out
        with torch.no_grad():
            return net.forward_seq(x)

    for i in range(args.warmup):
        step(xs[i % len(xs)])
    barrier()

    # ---- device-resident throughput
    if sampler:
        sampler.mark()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(xs[i % len(xs)])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the public API: pinned host frames in (copy overlapped with the previous step's kernels on a
    #      second stream), finest depth map back to pinned host memory, every step
    if train:
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        bufs = [torch.empty_like(xs[0]) for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(-2, args.steps):          # two untimed steps fill the pipeline
            if i == 0:
                barrier()
                e2.record()
            k = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[k])
                bufs[k].copy_(xs_host[i % len(xs_host)], non_blocking=True)
                ready[k].record(copy_stream)
            main.wait_event(ready[k])
            out = step(bufs[k])
            free[k].record(main)
            loss_host.copy_(out[0][0].mean(), non_blocking=True)
        e3.record()
        barrier()
        d2h_bytes = 4
    else:
        pipe = sb.pipeline.HostPipeline(net, tuple(xs_host[0].shape), dev, dtype=xs_host[0].dtype, stateless=not args.keep_state)
        for i in range(2):
            pipe.step(xs_host[i % len(xs_host)])
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(args.steps):
            depth_host = pipe.step(xs_host[i % len(xs_host)])
        pipe.join()              # the depth maps return on their own stream: the closing event waits for the last one
        e3.record()
        barrier()
        d2h_bytes = depth_host.numel() * 4
    ms_e2e = e2.elapsed_time(e3)
    h2d_bytes = xs_host[0].numel() * xs_host[0].element_size()

    # ---- per-kernel timing of the tensor-core blocks (roofline)
    eng = net.engine
    eng.timing = []
    for i in range(min(args.steps, 10)):
        with torch.no_grad():
            sb.functional.reset_net(net)
            net.forward_seq(xs[i % len(xs)])
    torch.cuda.synchronize()
    per_site = {}
    for name, a, b in eng.timing:
        per_site.setdefault(name, []).append(a.elapsed_time(b))
    eng.timing = None
    flop_scale = dict(getattr(eng, 'flop_scale', {}))      # folded decoder blocks execute (and are credited with) fewer taps

    # ---- the same loop with the other state policy, a few steps (reported beside the main number)
    alt_ms = None
    if not train:
        net.set_kernel_options(keep_state=not args.keep_state)
        for i in range(3):
            step(xs[i % len(xs)])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(min(args.steps, 10)):
            step(xs[i % len(xs)])
        a1.record()
        torch.cuda.synchronize()
        alt_ms = a0.elapsed_time(a1) / min(args.steps, 10)
        net.set_kernel_options(keep_state=bool(args.keep_state))

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])

    # ---- training sub-record (configs[2]/[3]) on the same ranks
    train_rec = None
    if not train and not args.no_train:
        del xs, pipe
        torch.cuda.empty_cache()
        tsteps, twarm = max(5, min(args.steps, 20)), 10
        ms_sync, ms_nosync, ncoll, tl = time_training(args, dev, world, rank, args.train_batch, tsteps, twarm, barrier)
        tt = torch.tensor([ms_sync, ms_nosync], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_sync, ms_nosync = float(tt[0]), float(tt[1])
        train_rec = {'metric': 'event-frames/sec', 'value': args.train_batch * T * world / (ms_sync / 1e3), 'unit': 'event-frames/s',
                     'n_gpus': world, 'steps': tsteps, 'warmup': twarm, 'ms_per_step': ms_sync, 'batch_per_gpu': args.train_batch,
                     'global_batch': args.train_batch * world, 'T': T,
                     'step': 'reset_net + forward_seq + Total_Loss (fused) + backward (surrogate BPTT, bf16 tensor-core dgrad/wgrad) + Adam',
                     'allreduce': {'bytes': 18148708 * 4, 'collectives_per_step': ncoll,
                                   'ms_per_step_without_allreduce': ms_nosync, 'exposed_ms': max(0.0, ms_sync - ms_nosync),
                                   'how': 'in-place NCCL all-reduce per block gradient issued inside the backward (reverse layer order), '
                                          'overlapped with the remaining dgrad/wgrad kernels; exposed = step time with - without it, max over ranks'},
                     'gpu_launches_per_step': tl}

    if rank == 0:
        peaks = measured_peaks()
        frames = B * T * world
        value = frames * args.steps / (ms / 1e3)
        e2e_val = frames * args.steps / (ms_e2e / 1e3)
        conv_sites = [k for k in per_site if k != 'heads']
        conv_ms = sum(statistics.median(per_site[k]) for k in conv_sites)      # median of 10 passes: an allocator growth inside one
        #                                                                         pass (cudaMalloc, ~10 ms) must not count as kernel time
        site_gflop = {k: MFLOP_PER_FRAME[k] * flop_scale.get(k, 1.0) / 1e3 * B * T for k in per_site if k in MFLOP_PER_FRAME}
        conv_gflop = sum(site_gflop[k] for k in conv_sites)
        ach = conv_gflop / conv_ms if conv_ms > 0 else 0.0      # GFLOP/ms == TFLOP/s
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        # the ncu captures behind traffic.json are of the headline configuration only (B=8 per GPU, T=5, 3 planes, inference)
        if os.path.isfile(tp) and B == 8 and T == 5 and args.planes == 3 and args.mode == 'infer':
            tj = json.load(open(tp))
            key = 'conv_i8_dram_bytes_per_step' + ('' if args.keep_state else '_stateless') + ('_fold' if args.fold else '')
            traffic = tj.get(key)
        line = {
            'metric': 'event-frames/sec', 'value': value, 'unit': 'event-frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8 x s8 -> s32 (tcgen05 kind::i8, %d weight digit planes), fp32 neuron state' % args.planes +
                                           ('; backward bf16 x bf16 -> f32 (tcgen05 kind::f16), fp32 surrogate scan' if args.mode == 'train' else ''),
            'data': 'synthetic', 'config': workload_config(args, B),
            'e2e': {'value': e2e_val, 'unit': 'event-frames/s', 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': ms_e2e / args.steps,
                    'host_frames': str(xs_host[0].dtype).replace('torch.', '') + ' ' + 'x'.join(str(int(v)) for v in xs_host[0].shape),
                    'api': ('forward_seq + backward + Adam on frames copied from pinned host memory on a second stream, loss scalar back to pinned memory' if train else
                            'stereospike_b200.pipeline.HostPipeline.step (pinned host frames -> H2D on a copy stream -> forward_seq -> pinned depth map on a copy-out stream)')},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_burst'], 'unit': 'TFLOP/s',
                         'frac': ach / peaks['bf16_burst'], 'traffic': traffic,
                         'frac_of_sustained_peak': ach / peaks['bf16_sustained'], 'peak_sustained': peaks['bf16_sustained'],
                         'kernel': 'conv_i8_kernel (all launches of the 13 fused blocks: bottom, conv1-4, bottleneck x4, deconv4-1); achieved, '
                                   'traffic and algorithmic work are summed over those launches',
                         'peak_source': peaks['source'] + ' BURST dense bf16 (MEASURED_PEAKS.json bf16_tflops: the kernels are timed one by one and the '
                                        'SM clock stays at its maximum); the kernel runs int8 MMAs, 3 digit planes = 1.5 bf16-equivalents per '
                                        'algorithmic FLOP, so 0.67 is the ceiling of this formulation at equal clocks',
                         'algorithmic_gflop_per_step': conv_gflop, 'kernel_ms_per_step': conv_ms,
                         'flops_basis': 'dense 2*M*N*K at reference geometry (SURVEY.md 8(a) row 5)' +
                                        ('; folded NNConvUpsampling blocks are credited with the taps they execute (flop_scale)' if flop_scale else ''),
                         'flop_scale': flop_scale or None,
                         'executed_int8_mac_factor': args.planes,
                         'per_block_ms': {k: round(statistics.median(v), 4) for k, v in per_site.items()},
                         'per_block_tflops': {k: round(site_gflop[k] / statistics.median(per_site[k]), 1) for k in site_gflop},
                         'per_block_frac': {k: round(site_gflop[k] / statistics.median(per_site[k]) / peaks['bf16_burst'], 3) for k in site_gflop}},
            'model_gflop_per_frame': TOTAL_GFLOP_PER_FRAME,
        }
        if alt_ms is not None:
            line['other_state_policy'] = {'keep_state': (not args.keep_state), 'ms_per_step': alt_ms,
                                          'value': B * T * world / (alt_ms / 1e3)}
        if train_rec is not None:
            line['train'] = train_rec
        if world == 1 and not train and not args.no_extras:
            line['timestep_sweep'] = timestep_sweep(args, dev)
            px = sj_cupy_proxy(args, dev, B, T)
            line['sj_cupy_proxy'] = {'value_fp32': px['fp32'], 'value_tf32_allowed': px['tf32'], 'unit': 'event-frames/s',
                                     'speedup_vs_fp32': value / px['fp32'], 'speedup_vs_tf32_allowed': value / px['tf32'],
                                     'what': 'PROXY (spikingjelly / cupy are not installable here): cuDNN convs over the flattened [T*B] batch '
                                             '+ one unfused per-neuron T-loop CUDA kernel per layer + elementwise gain / skip adds, fp32 NCHW, '
                                             f'forward, B={B}, T={T}, same GPU, same run'}
            line['analog_model'] = analog_model_record(dev)
        if world == 1 and not args.no_parity and not train:
            from tests._cases import parity_summary          # the oracle as the checker, outside every timed region
            line['parity'] = parity_summary(args.neuron, args.gain, args.tau, T=T, B=1, seed=0, planes=args.planes,
                                            fold=bool(args.fold))
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = cpu_oracle_rate(args.neuron, args.gain, args.tau, T, 1, 8)
            line['cpu_baseline'] = {'value': v, 'unit': 'event-frames/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'oracle/ref_model.py forward, B=1 x T={T} frames, median of 8 after 1 warm-up '
                                              f'({dt:.2f} s each), torch CPU fp32, {cores} threads'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='samples per GPU')
    ap.add_argument('--train-batch', type=int, default=16, help='samples per GPU of the training sub-record (configs[2]/[3])')
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'])
    ap.add_argument('--reference-device', default='cpu', choices=['cpu', 'cuda'],
                    help='--impl reference only: cuda = the oracle in PyTorch eager on the GPU (extra comparison point)')
    ap.add_argument('--T', type=int, default=5)
    ap.add_argument('--neuron', default='lif', choices=['if', 'lif', 'plif'])
    ap.add_argument('--gain', type=float, default=15.0)
    ap.add_argument('--tau', type=float, default=3.0)
    ap.add_argument('--planes', type=int, default=3)
    ap.add_argument('--kernel', default='umma', choices=['umma', 'simt'])
    ap.add_argument('--input-sets', type=int, default=4)
    ap.add_argument('--host-frames', default='u8', choices=['u8', 'f32'],
                    help='e2e arm: packed u8 count frames [T,B,H,W,4] (default) or the reference\'s fp32 [B,T,4,H,W] frames')
    ap.add_argument('--keep-state', type=int, default=0,
                    help='1 = write the final membrane potentials back after every step (stateful streaming); 0 = stateless serving')
    ap.add_argument('--fold', type=int, default=FOLD_DEFAULT, help='NNConvUpsampling blocks as folded 3x3 convs (fewer taps)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the timestep sweep and the SpikingJelly-cupy proxy (N=1 only)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
