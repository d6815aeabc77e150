"""CPU: oracle/events_ref.py against the reference's own event -> frame functions executed from /root/reference
(skipped where the reference is absent), and against a small committed known-answer case."""
import os
import sys
import types

import numpy as np
import pytest

from oracle import events_ref as er

REF = os.environ.get('STEREOSPIKE_REFERENCE', '/root/reference')


def _load_reference_utils():
    path = os.path.join(REF, 'datasets', 'MVSEC', 'utils.py')
    if not os.path.isfile(path):
        pytest.skip('reference not present')
    if 'h5py' not in sys.modules:                      # utils.py imports h5py at module level; it is not needed here
        sys.modules['h5py'] = types.ModuleType('h5py')
    import importlib.util
    spec = importlib.util.spec_from_file_location('_ref_mvsec_utils', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('nfpdm', [1, 5])
def test_cumulate_matches_reference(nfpdm):
    ref = _load_reference_utils()
    n_chunks = 3
    ev = er.synthetic_events(3000, n_chunks, nfpdm, seed=nfpdm)
    depth = np.zeros((n_chunks, 260, 346))
    ts = np.arange(n_chunks, dtype=np.float64) / 20 + ev[0, 2]
    want, _ = ref.mvsecCumulateSpikesIntoFrames(ev.copy(), depth, ts.copy(), num_frames_per_depth_map=nfpdm)
    got = er.cumulate_spikes_into_frames(ev, n_chunks, nfpdm)
    assert want.shape == got.shape and np.array_equal(want, got) and got.sum() > 2000


def test_rectify_matches_reference():
    ref = _load_reference_utils()
    rng = np.random.default_rng(3)
    xm = rng.uniform(-5, 350, (260, 346))
    ym = rng.uniform(-5, 264, (260, 346))
    ev = er.synthetic_events(2000, 2, 1, seed=9, raw=True)
    want = ref.mvsecRectifyEvents(ev.copy(), xm, ym)
    got = er.rectify_events(ev, xm, ym)
    assert np.array_equal(want, got) and 0 < len(got) < len(ev)


def test_known_answer():
    # three events in frame 0 (one OFF), one exactly on the 0.05 s boundary (dropped), one in frame 1
    ev = np.array([[10.7, 20.2, 0.0, 1], [10.1, 20.9, 0.01, 1], [3.0, 4.0, 0.02, -1], [7.0, 7.0, 0.05, 1], [1.0, 2.0, 0.07, 0]], dtype=np.float64)
    f = er.cumulate_spikes_into_frames(ev, 2, 1)
    assert f[0, 0, 0, 20, 10] == 1 and f[0, 0, 1, 4, 3] == 1      # the very first event sits at t = 0 and is not > start
    assert f[1, 0, 1, 2, 1] == 1 and f[:, :, :, 7, 7].sum() == 0 and f.sum() == 3
