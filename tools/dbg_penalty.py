import torch
import stereospike_b200 as sb
from oracle import loss_ref, ref_model as rm, sj_compat as sj
from stereospike_b200 import loss as sl
torch.manual_seed(4)
oracle = rm.SpikingUNet('lif', tau=3.0, multiply_factor=15.0)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0)
net.load_state_dict(oracle.state_dict())
net = net.cuda()
x = rm.synthetic_inputs(1, 2, 4, seed=8)
label = rm.synthetic_label(1, seed=9)
for only_pen in (True, False):
    for p in oracle.parameters(): p.grad = None
    sj.reset_net(oracle)
    d_ref, s_ref = oracle.forward_seq(x)
    lo = loss_ref.spike_penalization_loss(s_ref) if only_pen else loss_ref.total_loss(d_ref, label)
    lo.backward()
    ref = {k: (p.grad.clone() if p.grad is not None else None) for k, p in oracle.named_parameters()}
    for bwd in ('simt', 'umma'):
        net.set_kernel_options(bwd_impl=bwd)
        net.zero_grad()
        sb.functional.reset_net(net)
        d, s = net.forward_seq(x.cuda(), spikes_fp32=True)
        l = sl.SpikePenalization_Loss(s) if only_pen else sl.Total_Loss()(d, label.cuda())
        l.backward()
        out = []
        for k, p in net.named_parameters():
            if ref[k] is None or p.grad is None:
                out.append((k, None)); continue
            a, b = ref[k].flatten().double(), p.grad.cpu().flatten().double()
            out.append((k.replace('.0.weight', '').replace('.up.1.weight', ''), round(float((a @ b) / (a.norm() * b.norm() + 1e-30)), 5), f'{float(a.norm()):.2e}'))
        print('penalty only' if only_pen else 'depth loss', bwd, float(lo), float(l), out[:14])
print('spike sums', [float(a.sum()) for a in s_ref], [float(b.sum()) for b in s])
