"""How much of a training / inference step is host time?  Enqueue time (no sync) vs device time."""
import time
import torch
import stereospike_b200 as sb
from oracle import ref_model as rm
import bench
torch.manual_seed(0)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
opt = torch.optim.Adam(net.parameters(), lr=2e-4)
for mode, B in (('train', 16), ('infer', 8)):
    x = rm.synthetic_inputs(B, 5, 4, seed=0).cuda()
    label = rm.synthetic_label(B, seed=1).cuda()
    def step():
        sb.functional.reset_net(net)
        if mode == 'train':
            d = net.forward_seq(x)[0]
            bench.masked_l1(d, label).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        else:
            with torch.no_grad():
                net.forward_seq(x)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{mode}: host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, total {1e3 * (t2 - t0) / n:.2f} ms/step')
