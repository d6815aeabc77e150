"""Builds stereospike_b200/lib/libstereospike_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m stereospike_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lib', 'libstereospike_b200.so')
SOURCES = ('ss_api.cu', 'ss_simt.cu', 'ss_heads.cu', 'ss_conv_i8.cu', 'ss_bwd.cu', 'ss_events.cu', 'ss_wgrad_umma.cu', 'ss_loss.cu')
NVCC_FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '--threads', '0', '-I' + os.path.join(ROOT, 'include')]


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'stereospike_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: an instrumentation build next to the product library (e.g. build/timing/… with SS_ROLE_TIMING)."""
    if out is None and not force and not _stale():
        return OUT
    out = out or OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-o', out] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libstereospike_b200.so')
    return out


if __name__ == '__main__':
    if '--timing' in sys.argv:
        print(build(out=os.path.join(ROOT, 'build', 'timing', 'libstereospike_b200.so'), defines=('SS_ROLE_TIMING',)))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
