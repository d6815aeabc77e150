"""Bring-up helper: run every full-size layer geometry of the model as an isolated block case (own subprocess, timeout)
for a given T and B, to localise a failing configuration.  `python tools/gpu_t1_probe.py [T] [B]`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = {
    'bottom': dict(kind='conv', Cin=4, Cout=32, ks=5, Hin=260, Win=346, stride=1, pad=2, up=None),
    'conv1': dict(kind='conv', Cin=32, Cout=64, ks=5, Hin=260, Win=346, stride=2, pad=2, up=None),
    'conv2': dict(kind='conv', Cin=64, Cout=128, ks=5, Hin=130, Win=173, stride=2, pad=2, up=None),
    'conv3': dict(kind='conv', Cin=128, Cout=256, ks=5, Hin=65, Win=87, stride=2, pad=2, up=None),
    'conv4': dict(kind='conv', Cin=256, Cout=512, ks=5, Hin=33, Win=44, stride=2, pad=2, up=None),
    'bneck': dict(kind='conv', Cin=512, Cout=512, ks=3, Hin=17, Win=22, stride=1, pad=1, up=None),
    'deconv4': dict(kind='upconv', Cin=512, Cout=256, ks=5, Hin=17, Win=22, stride=1, pad=0, up=(33, 44)),
    'deconv3': dict(kind='upconv', Cin=256, Cout=128, ks=5, Hin=33, Win=44, stride=1, pad=0, up=(65, 87)),
    'deconv2': dict(kind='upconv', Cin=128, Cout=64, ks=5, Hin=65, Win=87, stride=1, pad=0, up=(130, 173)),
    'deconv1': dict(kind='upconv', Cin=64, Cout=32, ks=5, Hin=130, Win=173, stride=1, pad=0, up=(260, 346)),
}

if len(sys.argv) > 1 and sys.argv[1] == '--one':
    from tests._cases import block_case
    name, T, B = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    r = block_case(neuron=1, T=T, B=B, impl='umma', planes=3, resid=name.startswith('deconv'), gain=6.0, **LAYERS[name])
    print('RESULT ' + json.dumps({k: r[k] for k in ('max_dh_t0', 'spike_mismatch_outside_band', 'rate', 'ms')}))
else:
    T = sys.argv[1] if len(sys.argv) > 1 else '1'
    B = sys.argv[2] if len(sys.argv) > 2 else '1'
    for n in LAYERS:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', n, T, B], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
            print(f'[{n} T={T} B={B}]', lines[-1][7:] if lines else f'FAILED rc={p.returncode} :: {(p.stderr or p.stdout)[-300:]}', flush=True)
        except subprocess.TimeoutExpired:
            print(f'[{n}] TIMEOUT', flush=True)
