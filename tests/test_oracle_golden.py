"""oracle/ref_model.py against the committed fixtures (generated from the reference's own files by
oracle/make_golden.py).  Runs anywhere (no /root/reference needed)."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden as mg
from oracle import ref_model as rm
from oracle import sj_compat as sj


def build_oracle(variant, mono, multiply_factor, tau):
    return rm.SpikingUNet(variant, mono, surrogate_function=sj.ATan() if variant == 'if' else None,
                          tau=tau, multiply_factor=multiply_factor)


@pytest.mark.parametrize('name', sorted(mg.CASES))
def test_oracle_reproduces_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + '.npz'))
    got = mg.run_case(name, build=build_oracle)
    if not np.allclose(got['weight_checksum'], gold['weight_checksum'], rtol=0, atol=0):
        pytest.skip('torch default-init RNG differs from the build container; golden weights not reproducible')
    assert np.array_equal(got['x'], gold['x'])
    np.testing.assert_array_equal(got['depth1'], gold['depth1'])
    np.testing.assert_allclose(got['depth_sums'], gold['depth_sums'], rtol=1e-12)
    assert float(got['mde']) == float(gold['mde'])
    np.testing.assert_allclose(got['grad_l2'], gold['grad_l2'], rtol=1e-6)
    assert (gold['grad_l2'] > 0).all()
    if 'spk_nonzero' in gold.files:
        assert np.array_equal(got['spk_nonzero'], gold['spk_nonzero'])


def test_oracle_reproduces_analog_model_golden(golden_dir):
    """oracle/ann_ref.py against the fixture the reference's own network/ANN_models.py produced (eval forward, train forward +
    backward, BatchNorm running statistics)."""
    from oracle import ann_ref
    gold = np.load(os.path.join(golden_dir, 'ann_sigmoid.npz'))
    got = mg.run_ann_case(build=ann_ref.AnalogUNet)
    if not mg.weights_match_arrays(got['weight_checksum'], gold['weight_checksum']):
        pytest.skip('torch default-init RNG differs from the build container; golden weights not reproducible')
    assert np.array_equal(got['x'], gold['x'])
    np.testing.assert_array_equal(got['eval_depth1_sub'], gold['eval_depth1_sub'])
    np.testing.assert_allclose(got['eval_depth_sums'], gold['eval_depth_sums'], rtol=1e-12)
    assert float(got['eval_mde']) == float(gold['eval_mde'])
    np.testing.assert_allclose(got['train_depth_sums'], gold['train_depth_sums'], rtol=1e-9)
    np.testing.assert_allclose(got['train_loss'], gold['train_loss'], rtol=1e-7)
    np.testing.assert_allclose(got['grad_l2'], gold['grad_l2'], rtol=1e-5)
    np.testing.assert_allclose(got['running_mean_sum'], gold['running_mean_sum'], rtol=1e-9)
    assert (gold['grad_l2'] > 0).all() and list(got['grad_names']) == list(gold['grad_names'])
