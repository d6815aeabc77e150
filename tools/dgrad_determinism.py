"""Is ss_corr_bf16 (store / accumulate epilogues) bit-reproducible run to run?"""
import torch
from stereospike_b200 import ops
dev = torch.device('cuda')
for (Cin, Cout, ks, Hin, Win, stride, pad, T, B) in [(64, 64, 3, 17, 22, 1, 1, 5, 3), (512, 512, 3, 17, 22, 1, 1, 1, 3), (32, 64, 5, 37, 45, 2, 2, 2, 2)]:
    Hout, Wout = ops.conv_out_size(Hin, ks, stride, pad), ops.conv_out_size(Win, ks, stride, pad)
    geom = ops.BlockGeom('conv', Cin, Cout, ks, Hin, Win, Hout, Wout, stride, pad)
    gen = torch.Generator().manual_seed(3)
    w = ((torch.rand(Cout, Cin, ks, ks, generator=gen) * 2 - 1) / (Cin * ks * ks) ** 0.5).to(dev)
    gy = (torch.randn(T, B, Hout, Wout, Cout, generator=gen) * (torch.rand(T, B, Hout, Wout, Cout, generator=gen) < 0.5)).bfloat16().to(dev)
    plan = ops.DgradPlan(w, geom, dev)
    outs = []
    for rep in range(4):
        dst = torch.zeros((T, B, Hin, Win, Cin), dtype=torch.float32, device=dev)
        plan.run(gy, dst, T, B)
        torch.cuda.synchronize()
        outs.append(dst)
    for rep in range(1, 4):
        d = (outs[rep] - outs[0]).abs()
        print(f'{Cin}->{Cout} k{ks}s{stride}: rep {rep}: differing {int((d > 0).sum())} of {d.numel()}, max abs diff {float(d.max()):.3e} (scale {float(outs[0].abs().max()):.3e})')
