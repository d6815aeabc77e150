"""Single-step (T = 1) throughput with and without the batch-as-independent-steps launch: `python tools/t1_sweep.py [B ...]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stereospike_b200 as sb  # noqa: E402
from oracle import ref_model as rm  # noqa: E402  (synthetic-input recipe only)

Bs = [int(a) for a in sys.argv[1:]] or [16, 8, 4]
torch.manual_seed(0)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
for B in Bs:
    xs = [rm.synthetic_inputs(B, 1, 4, seed=700 + i).cuda() for i in range(2)]
    for on in (False, True, False, True):
        net.set_kernel_options(keep_state=False, batch_as_steps=on)
        with torch.no_grad():
            for i in range(5):
                sb.functional.reset_net(net)
                net.forward_seq(xs[i % 2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                sb.functional.reset_net(net)
                net.forward_seq(xs[i % 2])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f'B={B} T=1 batch_as_steps={on}: {ms:.4f} ms per call, {B / ms * 1e3:.0f} event-frames/s', flush=True)
