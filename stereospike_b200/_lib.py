"""ctypes binding of the C-ABI library (include/stereospike_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent, importing the
ops raises.  Build it with ``python -m stereospike_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('STEREOSPIKE_B200_LIB', os.path.join(_HERE, 'lib', 'libstereospike_b200.so'))   # env: instrumented builds

SS_NEURON_IF, SS_NEURON_LIF, SS_NEURON_PLIF = 0, 1, 2
SS_SURR_ATAN, SS_SURR_SIGMOID = 0, 1
SS_IN_U8_TBHWC, SS_IN_F32_BTCHW = 0, 1
SS_IMPL_AUTO, SS_IMPL_SIMT, SS_IMPL_UMMA = 0, 1, 2

# every symbol include/stereospike_b200.h declares (tests/test_cabi_symbols.py checks header == this == .so)
SYMBOLS = ('ss_events_accumulate', 'ss_events_pack', 'ss_conv_i8_fwd', 'ss_conv_i8_fwd_ex', 'ss_pack_digits_i8', 'ss_pack_digits_i8_rect', 'ss_pack_weights_folded', 'ss_conv_i8_rowbytes', 'ss_pack_weights_i8', 'ss_pack_events', 'ss_pack_events_c', 'ss_conv_neuron_fwd',
           'ss_heads_fwd', 'ss_neuron_fwd', 'ss_neuron_bwd', 'ss_neuron_bwd_ex', 'ss_conv_dgrad', 'ss_conv_wgrad', 'ss_heads_bwd',
           'ss_pack_weights_bf16', 'ss_corr_bf16', 'ss_conv_wgrad_bf16', 'ss_loss_fwd', 'ss_loss_bwd',
           'ss_abi_version', 'ss_last_error', 'ss_launch_count')


class BlockDesc(ctypes.Structure):
    _fields_ = [('T', ctypes.c_int32), ('B', ctypes.c_int32),
                ('Hin', ctypes.c_int32), ('Win', ctypes.c_int32), ('Cin', ctypes.c_int32),
                ('Hout', ctypes.c_int32), ('Wout', ctypes.c_int32), ('Cout', ctypes.c_int32),
                ('ks', ctypes.c_int32), ('stride', ctypes.c_int32), ('pad', ctypes.c_int32),
                ('upsample', ctypes.c_int32), ('neuron', ctypes.c_int32), ('planes', ctypes.c_int32),
                ('gain', ctypes.c_float), ('v_th', ctypes.c_float), ('v_reset', ctypes.c_float),
                ('tau', ctypes.c_float)]


class TileMaps(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('nclass', ctypes.c_int32), ('rl_n', ctypes.c_int32), ('transposed', ctypes.c_int32),
                ('ymap_out', ctypes.c_void_p), ('xmap_out', ctypes.c_void_p), ('rl_src', ctypes.c_void_p),
                ('rl_out', ctypes.c_void_p), ('rl_collive', ctypes.c_void_p), ('stats', ctypes.c_void_p),
                ('item_tab', ctypes.c_void_p), ('n_items', ctypes.c_int32), ('defer_wait', ctypes.c_int32),
                ('independent_steps', ctypes.c_int32)]


SS_TILES_PLAIN, SS_TILES_FOLDED, SS_TILES_ROW_LIST = 0, 1, 2


class ConvGeom(ctypes.Structure):
    _fields_ = [('T', ctypes.c_int32), ('B', ctypes.c_int32),
                ('Hin', ctypes.c_int32), ('Win', ctypes.c_int32), ('Cin', ctypes.c_int32),
                ('Hout', ctypes.c_int32), ('Wout', ctypes.c_int32), ('Cout', ctypes.c_int32),
                ('ks', ctypes.c_int32), ('in_layout', ctypes.c_int32), ('neuron', ctypes.c_int32),
                ('reserved0', ctypes.c_int32),
                ('gain', ctypes.c_float), ('v_th', ctypes.c_float), ('v_reset', ctypes.c_float),
                ('tau', ctypes.c_float),
                ('reserved1', ctypes.c_int32), ('reserved2', ctypes.c_int32)]


class CorrDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('T', 'B', 'Hg', 'Wg', 'Cg', 'Hv', 'Wv', 'Hdst', 'Wdst', 'Cdst', 'ks', 'pad',
                                               'nclass', 'out_mode', 'ntile', 'reserved')]


SS_CORR_STORE, SS_CORR_ACCUMULATE, SS_CORR_ATOMIC = 0, 1, 2


class HeadsArgs(ctypes.Structure):
    _fields_ = [('T', ctypes.c_int32), ('B', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32),
                ('gain', ctypes.c_float),
                ('C', ctypes.c_int32 * 4), ('Hs', ctypes.c_int32 * 4), ('Ws', ctypes.c_int32 * 4),
                ('acts', ctypes.c_void_p * 4), ('w', ctypes.c_void_p * 4), ('bias', ctypes.c_void_p * 4),
                ('ymap', ctypes.c_void_p * 4), ('xmap', ctypes.c_void_p * 4), ('taps', ctypes.c_void_p * 4),
                ('acts_sum', ctypes.c_void_p * 4)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library.  Raises LibraryMissing (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissing(
            f'{LIB_PATH} not found: the CUDA extension is not built. Run `python -m stereospike_b200.build`. '
            'There is no CPU or PyTorch fallback for the stereospike_b200 hot path.')
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    L.ss_conv_i8_fwd.argtypes = [ctypes.POINTER(BlockDesc)] + [vp] * 11
    L.ss_conv_i8_fwd.restype = ctypes.c_int
    L.ss_conv_i8_fwd_ex.argtypes = [ctypes.POINTER(BlockDesc), ctypes.POINTER(TileMaps)] + [vp] * 11
    L.ss_conv_i8_fwd_ex.restype = ctypes.c_int
    L.ss_pack_digits_i8.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    L.ss_pack_digits_i8.restype = ctypes.c_int
    L.ss_pack_digits_i8_rect.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.ss_pack_digits_i8_rect.restype = ctypes.c_int
    L.ss_pack_weights_folded.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ss_pack_weights_folded.restype = ctypes.c_int
    L.ss_events_accumulate.argtypes = [vp, i64, vp, vp, ctypes.c_double, vp, vp, i32, i32, i32, i32, vp, vp]
    L.ss_events_accumulate.restype = ctypes.c_int
    L.ss_events_pack.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    L.ss_events_pack.restype = ctypes.c_int
    L.ss_conv_i8_rowbytes.argtypes = [i32, i32]
    L.ss_conv_i8_rowbytes.restype = ctypes.c_int
    L.ss_pack_weights_i8.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.ss_pack_weights_i8.restype = ctypes.c_int
    L.ss_pack_events.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.ss_pack_events.restype = ctypes.c_int
    L.ss_pack_events_c.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    L.ss_pack_events_c.restype = ctypes.c_int
    L.ss_conv_neuron_fwd.argtypes = [ctypes.POINTER(ConvGeom)] + [vp] * 11
    L.ss_conv_neuron_fwd.restype = ctypes.c_int
    L.ss_heads_fwd.argtypes = [ctypes.POINTER(HeadsArgs), vp, vp, vp]
    L.ss_heads_fwd.restype = ctypes.c_int
    L.ss_heads_bwd.argtypes = [ctypes.POINTER(HeadsArgs), vp, vp * 4, vp * 4, vp * 4, vp * 4, i32, vp]
    L.ss_heads_bwd.restype = ctypes.c_int
    L.ss_neuron_fwd.argtypes = [i32, i64, i32, f32, f32, f32] + [vp] * 6
    L.ss_neuron_fwd.restype = ctypes.c_int
    L.ss_neuron_bwd.argtypes = [i32, i64, i32, i32, f32, f32, f32, f32, f32] + [vp] * 9
    L.ss_neuron_bwd.restype = ctypes.c_int
    L.ss_neuron_bwd_ex.argtypes = [i32, i64, i32, i32, f32, f32, f32, f32, f32] + [vp] * 10
    L.ss_neuron_bwd_ex.restype = ctypes.c_int
    L.ss_pack_weights_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.ss_pack_weights_bf16.restype = ctypes.c_int
    L.ss_corr_bf16.argtypes = [ctypes.POINTER(CorrDesc)] + [vp] * 6
    L.ss_corr_bf16.restype = ctypes.c_int
    L.ss_conv_wgrad_bf16.argtypes = [ctypes.POINTER(BlockDesc)] + [vp] * 4
    L.ss_conv_wgrad_bf16.restype = ctypes.c_int
    L.ss_loss_fwd.argtypes = [i32, i32, i32, i32, vp * 4, vp, vp, vp, vp]
    L.ss_loss_fwd.restype = ctypes.c_int
    L.ss_loss_bwd.argtypes = [i32, i32, i32, i32, vp * 4, vp, vp, vp, vp, vp, vp * 4, vp]
    L.ss_loss_bwd.restype = ctypes.c_int
    L.ss_conv_dgrad.argtypes = [ctypes.POINTER(ConvGeom)] + [vp] * 6
    L.ss_conv_dgrad.restype = ctypes.c_int
    L.ss_conv_wgrad.argtypes = [ctypes.POINTER(ConvGeom)] + [vp] * 6
    L.ss_conv_wgrad.restype = ctypes.c_int
    L.ss_abi_version.restype = ctypes.c_int
    L.ss_last_error.restype = ctypes.c_char_p
    L.ss_launch_count.restype = ctypes.c_int64
    for s in SYMBOLS:
        getattr(L, s)          # AttributeError here = header / library mismatch
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().ss_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def launch_count():
    return int(lib().ss_launch_count())
