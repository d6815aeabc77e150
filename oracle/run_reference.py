"""ORACLE -- test infrastructure only; never imported by the product path.

Loads the *unmodified* reference model files from /root/reference by path, with
``spikingjelly.clock_driven`` mapped to oracle/sj_compat.py.  Works around the reference's
HEAD breakage (network/__init__.py:2 imports a misspelt name, so ``import network`` raises)
by registering an empty ``network`` package whose __path__ points at the reference directory
and importing ``network.SNN_models`` directly.

/root/reference exists only in the build container, never on the GPU box: callers must
check ``available()`` and skip otherwise.
"""
import importlib
import os
import sys
import types

from . import sj_compat

REFERENCE_ROOT = os.environ.get('STEREOSPIKE_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'network', 'SNN_models.py'))


def load_reference_models():
    """Returns the reference's ``network.SNN_models`` module (executed from its own source)."""
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    sj_compat.install_shim()
    if 'network' not in sys.modules or not getattr(sys.modules['network'], '_oracle_pkg', False):
        pkg = types.ModuleType('network')
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'network')]
        pkg._oracle_pkg = True
        sys.modules['network'] = pkg
    return importlib.import_module('network.SNN_models')


def build_reference(variant='if', monocular=False, multiply_factor=5.0, tau=3.0):
    """Construct the reference class that corresponds to oracle.ref_model.SpikingUNet(variant, monocular)."""
    m = load_reference_models()
    if variant == 'if':
        assert not monocular
        return m.StereoSpike(surrogate_function=sj_compat.ATan(), multiply_factor=multiply_factor)
    cls = (m.fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike if monocular
           else m.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike)
    return cls(use_plif=(variant == 'plif'), tau=tau, multiply_factor=multiply_factor)


def build_reference_ann(activation_function=None):
    """The reference's StereoSpike_equivalentANN (network/ANN_models.py), executed from its own source."""
    load_reference_models()                       # installs the spikingjelly shim and the `network` package stub
    m = importlib.import_module('network.ANN_models')
    import torch.nn as nn
    return m.StereoSpike_equivalentANN(activation_function if activation_function is not None else nn.Sigmoid())
