#!/usr/bin/env python
"""Headline benchmark: event-frames/sec of the fused spiking U-Net forward (binocular, T=5, 260x346, batch 8 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of the whole path over one batch of synthetic event frames: B*T event-frames.
Rank 0 prints ONE JSON line (contract in the task statement / DESIGN.md section "Measurement").
  value      device-timed throughput, inputs resident in HBM (rotating input sets larger than L2)
  e2e        same metric through the public nn.Module call with pinned-HOST inputs: H2D of the step's frames and
             D2H of the finest depth map inside the timed region
  roofline   tcgen05 conv+neuron kernel: algorithmic FLOPs / CUDA-event time of those launches vs the measured
             sustained bf16 peak (MEASURED_PEAKS.json)
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H0, W0 = 260, 346
# forward conv FLOPs per event-frame of every tensor-core block (2*M*N*K at reference geometry)
MFLOP_PER_FRAME = {'bottom': 575.7, 'conv1': 2303.0, 'conv2': 2316.3, 'conv3': 2379.0, 'conv4': 2451.0,
                   'bottleneck.0.conv1': 1764.8, 'bottleneck.0.conv2': 1764.8, 'bottleneck.1.conv1': 1764.8,
                   'bottleneck.1.conv2': 1764.8, 'deconv4': 9515.8, 'deconv3': 9265.2, 'deconv2': 9211.9,
                   'deconv1': 9211.9, 'heads': 777.2}
TOTAL_GFLOP_PER_FRAME = sum(MFLOP_PER_FRAME.values()) / 1e3


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
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()][getattr(self, 'skip', 0):]
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
                'reasons': sorted(reasons), 'samples': len(sm)}


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
    t_all = time.perf_counter()
    import torch
    from oracle import ref_model as rm, sj_compat as sj
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = build_oracle(args.neuron, args.gain, args.tau)
    on_gpu = args.reference_device == 'cuda'
    sample_B = args.batch if on_gpu else 1
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
                         'sample': f'each step = B={sample_B} sample x T={args.T} frames of the same workload, {args.mode}, '
                                   f'torch {torch.__version__} ' + ('CUDA eager (cuDNN fp32, allow_tf32=False) -- the reference\'s own GPU '
                                   'execution model, NOT the CPU arm' if on_gpu else f'CPU fp32, {cores} threads')},
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
            'l2': f'rotating {args.input_sets} input sets per step; per-step activation stream (~2 GB) exceeds the 126 MB L2',
            'parallelism': f'replicas x{args.gpus} (batch shards, no data-path collective)'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import stereospike_b200 as sb
    from stereospike_b200 import _lib
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

    torch.manual_seed(0)
    if args.neuron == 'if':
        net = sb.StereoSpike(surrogate_function=sb.surrogate.ATan(), multiply_factor=args.gain)
    else:
        net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=args.neuron == 'plif', tau=args.tau,
                                                                              multiply_factor=args.gain)
    net = net.to(dev)
    net.set_kernel_options(impl=args.kernel, weight_planes=args.planes)
    B, T = args.batch, args.T
    xs_host = [rm.synthetic_inputs(B, T, 4, seed=100 + rank * 16 + i).pin_memory() for i in range(args.input_sets)]
    xs = [x.to(dev) for x in xs_host]

    train = args.mode == 'train'
    if train:
        label = rm.synthetic_label(B, seed=1).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=2e-4)
        sync = sb.parallel.GradientSynchronizer(net.parameters()) if world > 1 else None

    def step(x):
        sb.functional.reset_net(net)
        if train:
            out = net.forward_seq(x)
            masked_l1(out[0], label).backward()
            if sync is not None:
                sync.sync(local_samples=B, global_samples=B * world)     # NCCL all-reduce of the gradients
            opt.step()
            opt.zero_grad(set_to_none=True)
            return out
        with torch.no_grad():
            return net.forward_seq(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        # two static device buffers filled on a copy stream (as pipeline.HostPipeline does for inference): the transfer of
        # step i+1 overlaps the kernels of step i and no 115 MB tensor is allocated per step
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
            depth_host = out[0][0]
            loss_host.copy_(out[0][0].mean(), non_blocking=True)
        e3.record()
        barrier()
        d2h_bytes = 4
    else:
        pipe = sb.pipeline.HostPipeline(net, tuple(xs_host[0].shape), dev)
        for i in range(2):
            pipe.step(xs_host[i % len(xs_host)])
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(args.steps):
            depth_host = pipe.step(xs_host[i % len(xs_host)])
        e3.record()
        barrier()
        d2h_bytes = depth_host.numel() * 4
    ms_e2e = e2.elapsed_time(e3)

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

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])

    if rank == 0:
        peaks = measured_peaks()
        frames = B * T * world
        value = frames * args.steps / (ms / 1e3)
        e2e_val = frames * args.steps / (ms_e2e / 1e3)
        conv_sites = [k for k in per_site if k != 'heads']
        conv_ms = sum(statistics.mean(per_site[k]) for k in conv_sites)
        conv_gflop = sum(MFLOP_PER_FRAME[k] for k in conv_sites) / 1e3 * B * T
        ach = conv_gflop / conv_ms if conv_ms > 0 else 0.0      # GFLOP/ms == TFLOP/s
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        # the ncu capture behind traffic.json is of the headline configuration only (B=8 per GPU, T=5, 3 planes, inference)
        if os.path.isfile(tp) and B == 8 and T == 5 and args.planes == 3 and args.mode == 'infer':
            traffic = json.load(open(tp)).get('conv_i8_dram_bytes_per_step')
        line = {
            'metric': 'event-frames/sec', 'value': value, 'unit': 'event-frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8 x s8 -> s32 (tcgen05 kind::i8, %d weight digit planes), fp32 neuron state' % args.planes +
                                           ('; backward bf16 x bf16 -> f32 (tcgen05 kind::f16), fp32 surrogate scan' if args.mode == 'train' else ''),
            'data': 'synthetic', 'config': workload_config(args, B),
            'e2e': {'value': e2e_val, 'unit': 'event-frames/s', 'h2d_bytes_per_step': xs_host[0].numel() * 4,
                    'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': ms_e2e / args.steps,
                    'api': ('forward_seq + backward + Adam on frames copied from pinned host memory on a second stream, loss scalar back to pinned memory' if train else 'stereospike_b200.pipeline.HostPipeline.step (pinned fp32 frames -> forward_seq -> pinned depth map)')},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_sustained'], 'unit': 'TFLOP/s',
                         'frac': ach / peaks['bf16_sustained'], 'traffic': traffic,
                         'kernel': 'conv_i8_kernel (13 launches/step: bottom, conv1-4, bottleneck x4, deconv4-1); achieved, '
                                   'traffic and algorithmic work are summed over those 13 launches',
                         'peak_source': peaks['source'] + ' sustained dense bf16 (MEASURED_PEAKS.json; the kernel runs int8 MMAs, '
                                        '3 digit planes = 1.5 bf16-equivalents per algorithmic FLOP)',
                         'algorithmic_gflop_per_step': conv_gflop, 'kernel_ms_per_step': conv_ms,
                         'executed_int8_mac_factor': args.planes,
                         'per_block_ms': {k: round(statistics.mean(v), 4) for k, v in per_site.items()},
                         'per_block_tflops': {k: round(MFLOP_PER_FRAME[k] / 1e3 * B * T / statistics.mean(v), 1)
                                              for k, v in per_site.items() if k in MFLOP_PER_FRAME}},
            'model_gflop_per_frame': TOTAL_GFLOP_PER_FRAME,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = cpu_oracle_rate(args.neuron, args.gain, args.tau, T, 1, 8)
            line['cpu_baseline'] = {'value': v, 'unit': 'event-frames/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'oracle/ref_model.py forward, B=1 x T={T} frames, median of 8 after 1 warm-up '
                                              f'({dt:.2f} s each), torch CPU fp32, {cores} threads'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='samples per GPU')
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
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
