"""Where does the weight-gradient kernel spend its time?  Instrumentation build (python -m stereospike_b200.build --timing):
SS_WG_DBG bit 1 = no MMA issue, 2 = no x-patch staging, 4 = no g-tile staging; each variant runs in its own interpreter:
    for d in 0 1 2 4 6 7; do SS_WG_DBG=$d STEREOSPIKE_B200_LIB=build/timing/libstereospike_b200.so PYTHONPATH=. python tools/wgrad_probe.py; done
"""
import os
import torch
from stereospike_b200 import ops

dev = torch.device('cuda')
T, B = 5, 16
CASES = [('deconv4', 'upconv', 512, 256, 5, 17, 22, 1, 0, (33, 44)), ('deconv3', 'upconv', 256, 128, 5, 33, 44, 1, 0, (65, 87)),
         ('deconv2', 'upconv', 128, 64, 5, 65, 87, 1, 0, (130, 173)), ('deconv1', 'upconv', 64, 32, 5, 130, 173, 1, 0, (260, 346)),
         ('bottleneck', 'conv', 512, 512, 3, 17, 22, 1, 1, None), ('conv3', 'conv', 128, 256, 5, 65, 87, 2, 2, None)]
for name, kind, Cin, Cout, ks, Hin, Win, stride, pad, up in CASES:
    if kind == 'conv':
        Hout, Wout = ops.conv_out_size(Hin, ks, stride, pad), ops.conv_out_size(Win, ks, stride, pad)
        geom = ops.BlockGeom('conv', Cin, Cout, ks, Hin, Win, Hout, Wout, stride, pad)
    else:
        geom = ops.BlockGeom('upconv', Cin, Cout, ks, Hin, Win, up[0], up[1])
    x = (torch.rand(T, B, Hin, Win, Cin, device=dev) < 0.15).to(torch.uint8)
    g = torch.randn(T, B, geom.Hout, geom.Wout, Cout, device=dev).bfloat16()
    for _ in range(2):
        ops.conv_wgrad_bf16(x, g, geom, T, B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_wgrad_bf16(x, g, geom, T, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flop = 2.0 * T * B * geom.Hout * geom.Wout * Cout * Cin * ks * ks
    import ctypes
    import numpy as np
    from stereospike_b200 import _lib
    L = _lib.lib()
    if hasattr(L, 'ss_wg_debug_read'):
        buf = np.zeros(24, dtype=np.uint64)
        L.ss_wg_debug_read.argtypes = [ctypes.c_void_p]
        if L.ss_wg_debug_read(buf.ctypes.data) == 0:
            d = buf.reshape(3, 8).astype(float)
            print(f'    CTA 0: x producer {d[0, 7]:.0f} cycles: wait stage {100 * d[0, 0] / d[0, 7]:.0f}%, convert+store {100 * d[0, 1] / d[0, 7]:.0f}%, '
                  f'wait last MMA {100 * d[0, 2] / d[0, 7]:.0f}% | MMA thread {d[2, 7]:.0f} cycles: wait x {100 * d[2, 0] / d[2, 7]:.0f}%, '
                  f'wait g {100 * d[2, 1] / d[2, 7]:.0f}%, issue {100 * d[2, 2] / d[2, 7]:.0f}%')
    print(f'dbg={os.environ.get("SS_WG_DBG", "0")} n64={os.environ.get("SS_WGRAD_N64", "1")} {name:10s} {ms:7.3f} ms  {flop / ms / 1e9:7.1f} TFLOP/s')
