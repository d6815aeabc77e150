import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stereospike_b200 as sb
from stereospike_b200 import ops
from oracle import ref_model as rm
orig = ops.conv_i8_fwd
def traced(x, g, *a, **k):
    print('conv_i8_fwd', g.kind, g.Cin, g.Cout, g.ks, g.Hin, g.Win, '->', g.Hout, g.Wout, 'T', k['T'], 'B', k['B'], 'v_in', k.get('v_in') is not None, 'neuron', k['neuron'], flush=True)
    r = orig(x, g, *a, **k)
    torch.cuda.synchronize()
    return r
ops.conv_i8_fwd = traced
torch.manual_seed(3)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
x = rm.synthetic_inputs(1, 3, 4, seed=9).cuda()
with torch.no_grad():
    sb.functional.reset_net(net)
    d_seq, s_seq = net.forward_seq(x)
    print('forward_seq ok', flush=True)
    sb.functional.reset_net(net)
    for t in range(3):
        d_it, s_it = net(x[:, t:])
        print('step', t, 'ok', flush=True)
