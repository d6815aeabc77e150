"""The restated U-Net (oracle/ref_model.py) must agree BIT-FOR-BIT, forward and backward, with the
reference's own files executed by path (only possible in the build container)."""
import pytest
import torch

from oracle import ref_model as rm
from oracle import run_reference as rr
from oracle import sj_compat as sj
from oracle.make_golden import simple_loss

pytestmark = pytest.mark.skipif(not rr.available(), reason='/root/reference not present (GPU box)')


@pytest.mark.parametrize('variant,mono,gain', [('if', False, 5.0), ('lif', False, 15.0), ('plif', True, 15.0)])
def test_bit_identical_forward_backward(variant, mono, gain):
    torch.manual_seed(7)
    ref = rr.build_reference(variant, mono, multiply_factor=gain, tau=3.0)
    mine = rm.SpikingUNet(variant, mono, surrogate_function=sj.ATan() if variant == 'if' else None,
                          tau=3.0, multiply_factor=gain)
    assert sorted(ref.state_dict()) == sorted(mine.state_dict())
    assert ref.count_trainable_params() == sum(p.numel() for p in mine.parameters())
    mine.load_state_dict(ref.state_dict())
    x = rm.synthetic_inputs(1, 2, 2 if mono else 4, seed=3)
    label = rm.synthetic_label(1, seed=4)
    sj.reset_net(ref)
    sj.reset_net(mine)
    for t in range(2):
        o_ref = ref(x[:, t:t + 1])
    o_me = mine.forward_seq(x)
    d_ref, d_me = (o_ref, o_me) if mono else (o_ref[0], o_me[0])
    for a, b in zip(d_ref, d_me):
        assert torch.equal(a, b)
    if not mono:
        for a, b in zip(o_ref[1], o_me[1]):
            assert torch.equal(a, b)
        assert 0.01 < float(o_me[1][0].count_nonzero()) / o_me[1][0].numel() < 0.9
    simple_loss(d_ref, label).backward()
    simple_loss(d_me, label).backward()
    g_ref = dict(ref.named_parameters())
    nz = 0
    for k, p in mine.named_parameters():
        assert torch.equal(p.grad, g_ref[k].grad), k
        nz += int(p.grad.abs().sum() > 0)
    assert nz == len(g_ref)          # a live network: every parameter receives gradient


def test_binocular_param_count():
    assert sum(p.numel() for p in rm.SpikingUNet('if').parameters()) == 18148708


@pytest.mark.parametrize('train', [False, True])
def test_analog_comparison_model_bit_identical(train):
    """oracle/ann_ref.py against network/ANN_models.py executed by path: same keys, same depths, same gradients."""
    from oracle import ann_ref
    torch.manual_seed(11)
    ref = rr.build_reference_ann()
    mine = ann_ref.AnalogUNet()
    assert sorted(ref.state_dict()) == sorted(mine.state_dict())
    ann_ref.randomize_batchnorm(ref, seed=5)
    mine.load_state_dict(ref.state_dict())
    ref.train(train)
    mine.train(train)
    x = rm.synthetic_inputs(2 if train else 1, 1, 4, seed=3)
    label = rm.synthetic_label(2 if train else 1, seed=4)
    sj.reset_net(ref)
    d_ref, d_me = ref(x), mine(x)
    for a, b in zip(d_ref, d_me):
        assert torch.equal(a, b)
    simple_loss(d_ref, label).backward()
    simple_loss(d_me, label).backward()
    g_ref = dict(ref.named_parameters())
    for k, p in mine.named_parameters():
        assert torch.equal(p.grad, g_ref[k].grad), k
    if train:
        for (k, a), (_, b) in zip(ref.named_buffers(), mine.named_buffers()):
            assert torch.equal(a, b), k
