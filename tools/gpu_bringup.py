"""Bring-up harness (GPU box): runs each kernel case in its own subprocess (a trap / illegal instruction poisons the
CUDA context) with a timeout, and prints compact error metrics.  `python tools/gpu_bringup.py [case ...]`."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


from tests._cases import block_case, model_case  # noqa: E402


CASES = {
    # name: kwargs
    'simt_conv_s2_if': dict(kind='conv', Cin=32, Cout=64, ks=5, Hin=20, Win=27, stride=2, pad=2, up=None, neuron=0, T=3, B=2, impl='simt', planes=0, resid=False),
    'simt_up_lif_res': dict(kind='upconv', Cin=64, Cout=32, ks=5, Hin=9, Win=11, stride=1, pad=0, up=(19, 23), neuron=1, T=3, B=2, impl='simt', planes=0, resid=True),
    'simt_3x3_plif': dict(kind='conv', Cin=64, Cout=64, ks=3, Hin=7, Win=9, stride=1, pad=1, up=None, neuron=2, T=4, B=1, impl='simt', planes=0, resid=True),
    'umma_conv64_if_T1': dict(kind='conv', Cin=64, Cout=128, ks=3, Hin=12, Win=13, stride=1, pad=1, up=None, neuron=0, T=1, B=1, impl='umma', planes=2, resid=False),
    'umma_conv64_if_p3': dict(kind='conv', Cin=64, Cout=128, ks=3, Hin=12, Win=13, stride=1, pad=1, up=None, neuron=0, T=3, B=2, impl='umma', planes=3, resid=False),
    'umma_conv32_s2_if': dict(kind='conv', Cin=32, Cout=64, ks=5, Hin=20, Win=27, stride=2, pad=2, up=None, neuron=0, T=3, B=2, impl='umma', planes=3, resid=False),
    'umma_up_lif_res_n32': dict(kind='upconv', Cin=64, Cout=32, ks=5, Hin=9, Win=11, stride=1, pad=0, up=(19, 23), neuron=1, T=3, B=2, impl='umma', planes=3, resid=True),
    'umma_3x3_512_plif': dict(kind='conv', Cin=512, Cout=512, ks=3, Hin=17, Win=22, stride=1, pad=1, up=None, neuron=2, T=2, B=1, impl='umma', planes=2, resid=True),
    'model_if_simt': dict(model=True, variant='if', mono=False, gain=5.0, T=2, B=1, impl='simt', planes=3),
    'model_if_umma': dict(model=True, variant='if', mono=False, gain=5.0, T=2, B=1, impl='umma', planes=3),
    'model_lif_umma': dict(model=True, variant='lif', mono=False, gain=15.0, T=3, B=1, impl='umma', planes=3),
    'model_plif_mono_umma': dict(model=True, variant='plif', mono=True, gain=15.0, T=2, B=2, impl='umma', planes=3),
    'model_if_simt_bwd': dict(model=True, variant='if', mono=False, gain=5.0, T=2, B=1, impl='simt', planes=3, backward=True),
    'model_plif_umma_bwd': dict(model=True, variant='plif', mono=False, gain=15.0, T=2, B=1, impl='umma', planes=3, backward=True),
    'umma_conv4_s2': dict(kind='conv', Cin=256, Cout=512, ks=5, Hin=33, Win=44, stride=2, pad=2, up=None, neuron=0, T=2, B=2, impl='umma', planes=3, resid=False),
    'umma_big_up': dict(kind='upconv', Cin=128, Cout=64, ks=5, Hin=33, Win=44, stride=1, pad=0, up=(65, 87), neuron=0, T=2, B=2, impl='umma', planes=3, resid=True),
}


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--one':
        name = sys.argv[2]
        kw = dict(CASES[name])
        if kw.pop('model', False):
            print('RESULT ' + json.dumps(model_case(**kw)))
        else:
            print('RESULT ' + json.dumps(block_case(**kw)))
        return
    names = sys.argv[1:] or list(CASES)
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', n], capture_output=True, text=True, timeout=600)
            lines = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
            if lines:
                print(f'[{n}] {lines[-1][7:]}')
            else:
                print(f'[{n}] FAILED rc={p.returncode} :: {(p.stderr or p.stdout)[-600:]}')
        except subprocess.TimeoutExpired:
            print(f'[{n}] TIMEOUT')
        print(f'    ({time.time() - t0:.1f}s)', flush=True)


if __name__ == '__main__':
    main()
