"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports exactly the symbols the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from stereospike_b200 import build, _lib
    build.build()
    return _lib.lib()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'stereospike_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ss_[a-z0-9_]+)\s*\(', src)))


def test_header_matches_binding_and_library(lib):
    from stereospike_b200 import _lib
    hs = header_symbols()
    assert hs == sorted(_lib.SYMBOLS), (hs, sorted(_lib.SYMBOLS))
    for s in hs:
        assert hasattr(lib, s), s


def test_abi_version_and_launch_counter(lib):
    assert lib.ss_abi_version() == 3
    assert lib.ss_launch_count() >= 0
    assert isinstance(lib.ss_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Entry points validate before touching the device: null / malformed arguments return SS_EINVAL (-1)."""
    from stereospike_b200 import _lib
    g = _lib.ConvGeom(T=1, B=1, Hin=4, Win=4, Cin=8, Hout=4, Wout=4, Cout=32, ks=3, in_layout=0, neuron=0, reserved0=0,
                      gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0, reserved1=0, reserved2=0)
    rc = lib.ss_conv_neuron_fwd(ctypes.byref(g), None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b'null' in lib.ss_last_error()
    d = _lib.BlockDesc(T=1, B=1, Hin=4, Win=4, Cin=32, Hout=4, Wout=4, Cout=32, ks=3, stride=1, pad=1, upsample=0, neuron=0,
                       planes=3, gain=1.0, v_th=1.0, v_reset=0.0, tau=2.0)
    rc = lib.ss_conv_i8_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b'null' in lib.ss_last_error()
    assert lib.ss_pack_weights_i8(None, 0, 0, 0, 0, None, None, None, None) == -1
    assert lib.ss_pack_events(None, 1, 1, 4, 2, 2, None, None, None) == -1
    assert lib.ss_conv_i8_rowbytes(512, 3) == 64 and lib.ss_conv_i8_rowbytes(64, 5) == 32
    assert lib.ss_neuron_bwd(1, 1, 0, 0, 2.0, 1.0, 1.0, 0.0, 2.0, None, None, None, None, None, None, None, None, None) == -1


def test_struct_layout_matches_header():
    from stereospike_b200 import _lib
    assert ctypes.sizeof(_lib.ConvGeom) == 18 * 4
    assert ctypes.sizeof(_lib.BlockDesc) == 18 * 4
    # ss_heads_args: 4 int32 + float + 12 int32 (+4 pad) + 28 pointers
    assert ctypes.sizeof(_lib.HeadsArgs) == 72 + 28 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from stereospike_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.LibraryMissing):
        _lib.lib()


def test_no_cpu_fallback():
    import torch
    import stereospike_b200 as sb
    net = sb.StereoSpike(multiply_factor=5.0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net(torch.zeros(1, 1, 4, 260, 346))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'stereospike_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn
