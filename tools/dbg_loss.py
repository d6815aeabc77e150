import torch
import stereospike_b200 as sb
from oracle import loss_ref, ref_model as rm
from stereospike_b200 import loss as sl
torch.manual_seed(3)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
x = rm.synthetic_inputs(1, 2, 4, seed=6).cuda()
label = rm.synthetic_label(1, seed=7).cuda()
with torch.no_grad():
    pred, _ = net.forward_seq(x)
pred = [p.detach().clone() for p in pred]
def grads(fn, tf32, dev='cuda'):
    torch.backends.cudnn.allow_tf32 = tf32
    ps = [p.to(dev).clone().requires_grad_(True) for p in pred]
    l = fn(ps, label.to(dev))
    l.backward()
    return float(l), [p.grad.cpu() for p in ps]
lf, gf = grads(lambda p, l: sl.Total_Loss()(p, l), True)
for name, tf32, dev in (('oracle cuda tf32', True, 'cuda'), ('oracle cuda fp32', False, 'cuda'), ('oracle cpu', False, 'cpu')):
    lo, go = grads(lambda p, l: loss_ref.total_loss(p, l), tf32, dev)
    print(name, 'loss', lo, 'fused', lf, [float((a - b).abs().max() / b.abs().max()) for a, b in zip(gf, go)],
          [int(((a - b).abs() > 1e-3 * b.abs().max()).sum()) for a, b in zip(gf, go)])
print('pred stats', [(float(p.min()), float(p.max()), float((p[..., 1:] == p[..., :-1]).float().mean())) for p in pred])
