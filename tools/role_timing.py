"""Per-role cycle breakdown of conv_i8_kernel from the instrumentation build (run on the GPU box):
    nvcc ... -DSS_ROLE_TIMING -o build/timing/libstereospike_b200.so   
    STEREOSPIKE_B200_LIB=build/timing/libstereospike_b200.so PYTHONPATH=. python tools/role_timing.py
"""
import ctypes
import numpy as np
import torch
from stereospike_b200 import _lib, ops

L = _lib.lib()
L.ss_debug_read.argtypes = [ctypes.c_void_p]
NAMES = {0: ('producer', ['geometry', 'wait free stage', 'copy issue']),
         1: ('mma thread 0', ['wait patch', 'wait token', 'issue', 'wait free slot', 'wait weights']),
         2: ('mma thread 1', ['wait patch', 'wait token', 'issue', 'wait free slot', 'wait weights']),
         3: ('epilogue', ['wait accumulator'])}


def run(name, kind, Cin, Cout, ks, Hin, Win, stride, pad, up, T, B):
    dev = torch.device('cuda')
    if kind == 'conv':
        Hout, Wout = ops.conv_out_size(Hin, ks, stride, pad), ops.conv_out_size(Win, ks, stride, pad)
        geom = ops.BlockGeom('conv', Cin, Cout, ks, Hin, Win, Hout, Wout, stride, pad)
    else:
        geom = ops.BlockGeom('upconv', Cin, Cout, ks, Hin, Win, up[0], up[1])
    x = (torch.rand(T, B, Hin, Win, Cin, device=dev) < 0.1).to(torch.uint8)
    w = (torch.rand(Cout, Cin, ks, ks, device=dev) * 2 - 1) / (Cin * ks * ks) ** 0.5
    q, sc, _ = ops.pack_weights_i8(w, 3, cin_pad=4 if Cin <= 4 else None)
    for _ in range(2):
        ops.conv_i8_fwd(x, geom, q, sc, T=T, B=B, neuron=1, gain=15.0, v_th=1.0, v_reset=0.0, tau=3.0, want_v_out=True,
                        cin=Cin)
    torch.cuda.synchronize()
    buf = np.zeros(148 * 4 * 8, dtype=np.uint64)
    assert L.ss_debug_read(buf.ctypes.data) == 0
    d = buf.reshape(148, 4, 8).astype(np.float64)
    import os
    if os.environ.get('SS_PAIR', '1') != '0':
        d = d[0::2]           # CTA pairs: only rank 0 (even CTAs) issues MMAs; its lines describe the pair
    total = d[:, 0, 7].mean()
    print(f'== {name}: T={T} B={B}  kernel ~{total:.0f} cycles per CTA')
    for role, (rn, fields) in NAMES.items():
        tot = d[:, role, 7].mean()
        parts = ', '.join(f'{f} {100 * d[:, role, i].mean() / max(tot, 1):.0f}%' for i, f in enumerate(fields))
        print(f'   {rn:13s} ({tot:9.0f} cycles): {parts}')


for T in (5,):
    run('bottom', 'conv', 4, 32, 5, 260, 346, 1, 2, None, T, 16)
    run('conv1', 'conv', 32, 64, 5, 260, 346, 2, 2, None, T, 16)
    run('conv3', 'conv', 128, 256, 5, 65, 87, 2, 2, None, T, 16)
    run('deconv1', 'upconv', 64, 32, 5, 130, 173, 1, 0, (260, 346), T, 16)
    run('deconv4', 'upconv', 512, 256, 5, 17, 22, 1, 0, (33, 44), T, 16)
