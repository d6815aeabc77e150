"""Pins oracle/loss_ref.py by executing the reference's own network/loss.py and network/metrics.py (pure PyTorch) on the
same inputs, values and gradients.  Needs /root/reference, so it only runs in the build container."""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = '/root/reference/network'
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'loss.py')), reason='reference checkout not present')


def _load(name):
    if 'matplotlib' not in sys.modules:          # metrics.py imports pyplot for its plotting helpers only
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mp = types.ModuleType('matplotlib')
            mp.pyplot = types.ModuleType('matplotlib.pyplot')
            sys.modules['matplotlib'] = mp
            sys.modules['matplotlib.pyplot'] = mp.pyplot
    spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(REF, name + '.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _inputs(seed, B=2, H=26, W=35, nan_frac=0.2):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, H, W, generator=g) * 10
    gt[torch.rand(B, 1, H, W, generator=g) < nan_frac] = float('nan')
    preds = [torch.rand(B, 1, H, W, generator=g) * 10 for _ in range(4)]
    return preds, gt


@pytest.mark.parametrize('seed', [0, 1])
def test_total_loss_and_gradients_match_the_reference(seed):
    from oracle import loss_ref
    ref = _load('loss')
    preds, gt = _inputs(seed)
    pa = [p.clone().requires_grad_(True) for p in preds]
    pb = [p.clone().requires_grad_(True) for p in preds]
    la = ref.Total_Loss(alpha=0.5, scale_weights=(1., 0.5, 2., 1.))(pa, gt)
    lb = loss_ref.total_loss(pb, gt, alpha=0.5, scale_weights=(1., 0.5, 2., 1.))
    assert torch.equal(la, lb)
    la.backward()
    lb.backward()
    for a, b in zip(pa, pb):
        assert torch.equal(a.grad, b.grad)


def test_single_terms_and_metric_match_the_reference():
    from oracle import loss_ref
    ref, met = _load('loss'), _load('metrics')
    preds, gt = _inputs(3)
    assert torch.equal(ref.ScaleInvariant_Loss(preds[0], gt), loss_ref.scale_invariant_loss(preds[0], gt))
    assert torch.equal(ref.GradientMatching_Loss(preds[1], gt), loss_ref.gradient_matching_loss(preds[1], gt))
    assert torch.equal(met.MeanDepthError(preds[2], gt), loss_ref.mean_depth_error(preds[2], gt))
    spikes = [(torch.rand(2, 8, 5, 7) < 0.3).float() * 2 for _ in range(3)]
    assert torch.equal(ref.SpikePenalization_Loss(spikes), loss_ref.spike_penalization_loss(spikes))
