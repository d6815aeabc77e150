import torch
import stereospike_b200 as sb
from oracle import loss_ref, ref_model as rm
torch.manual_seed(3)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
x = rm.synthetic_inputs(1, 2, 4, seed=6).cuda()
label = rm.synthetic_label(1, seed=7).cuda()
def run(bwd):
    net.set_kernel_options(bwd_impl=bwd)
    net.zero_grad()
    sb.functional.reset_net(net)
    pred, _ = net.forward_seq(x)
    loss_ref.total_loss(pred, label).backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in net.named_parameters()}
a, b, c, d = run('umma'), run('umma'), run('simt'), run('simt')
for k in a:
    m = float(c[k].abs().max()) + 1e-30
    print(f'{k:34s} umma-umma {float((a[k]-b[k]).abs().max())/m:.2e}  simt-simt {float((c[k]-d[k]).abs().max())/m:.2e}  umma-simt {float((a[k]-c[k]).abs().max())/m:.2e}')
