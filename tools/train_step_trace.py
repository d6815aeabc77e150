"""Per-step device times of the training step (B = 16, T = 5): `python tools/train_step_trace.py [steps]` -- shows whether a slow
run is uniformly slow or a few slow steps, and what the caching allocator did meanwhile."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stereospike_b200 as sb  # noqa: E402
from oracle import ref_model as rm  # noqa: E402  (synthetic-input recipe only)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, T = 16, 5
torch.manual_seed(0)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
opt = torch.optim.Adam(net.parameters(), lr=2e-4)
crit = sb.loss.Total_Loss(alpha=0.5)
xs = [rm.synthetic_inputs(B, T, 4, seed=500 + i).cuda() for i in range(2)]
label = rm.synthetic_label(B, seed=600).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
stats0 = None
for i in range(10 + steps):
    if i == 10:
        torch.cuda.synchronize()
        stats0 = torch.cuda.memory_stats()
        ev[0].record()
    sb.functional.reset_net(net)
    pred, _ = net.forward_seq(xs[i % 2])
    loss = crit(pred, label)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    if i >= 10:
        ev[i - 10 + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
st = torch.cuda.memory_stats()
print('per-step ms: min %.2f median %.2f mean %.2f max %.2f' % (min(ms), statistics.median(ms), statistics.mean(ms), max(ms)))
print(' '.join('%.1f' % m for m in ms))
for k in ('num_alloc_retries', 'num_device_alloc', 'num_device_free', 'reserved_bytes.all.peak', 'allocated_bytes.all.peak'):
    print(k, st.get(k), '(+%s during the timed steps)' % (st.get(k, 0) - stats0.get(k, 0)))
