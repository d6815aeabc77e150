"""Single-sample latency of stateless inference, eager launches vs CUDA-graph replay (run on the GPU box)."""
import json
import sys
import torch
import stereospike_b200 as sb
from oracle import ref_model as rm
from stereospike_b200.pipeline import GraphedInference

torch.manual_seed(0)
net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
for B, T in ((1, 1), (1, 5), (8, 5)):
    x = rm.synthetic_inputs(B, T, 4, seed=0).cuda()
    run = GraphedInference(net, tuple(x.shape))

    def eager():
        sb.functional.reset_net(net)
        with torch.no_grad():
            net.forward_seq(x)
    res = {}
    for name, fn in (('eager', eager), ('graph', lambda: run(x))):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / n
    print(json.dumps({'B': B, 'T': T, 'eager_ms': round(res['eager'], 4), 'graph_ms': round(res['graph'], 4),
                      'event_frames_per_s_graph': round(B * T / res['graph'] * 1e3, 1)}), flush=True)
