"""stereospike_b200 -- B200-native (sm_100a) implementation of the StereoSpike spiking encoder-decoder hot path.

Public surface mirrors the reference (urancon/StereoSpike ``network`` package + the SpikingJelly pieces it uses):
``stereospike_b200.models`` (StereoSpike, fromZero_* classes), ``.blocks`` (MultiplyBy, NNConvUpsampling,
SEWResBlock), ``.ann`` (the analog comparison model of network/ANN_models.py), ``.neuron`` / ``.surrogate`` / ``.functional`` (SpikingJelly replacements), ``.loss`` (network/loss.py +
MeanDepthError), ``.events`` (event stream -> input frames).
"""
from . import ann, blocks, events, functional, loss, models, neuron, parallel, pipeline, surrogate  # noqa: F401
from .blocks import MultiplyBy, NNConvUpsampling, SEWResBlock  # noqa: F401
from .models import (NeuromorphicNet, StereoSpike,  # noqa: F401
                     fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike,
                     fromZero_feedforward_multiscale_tempo_monocular_SpikeFlowNetLike)

__version__ = '0.1.0'
