"""The analog comparison model (stereospike_b200.ann, network/ANN_models.py upstream) on the GPU against the CPU oracle
(oracle/ann_ref.py, pinned bit for bit against the reference's own file) and the committed fixture the reference produced."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_MDE = 1e-3          # north_star tolerance on the depth metric
TOL_DEPTH = 2e-4        # fp32 convolutions in another summation order, 13 layers deep: relative to the largest |depth|


def _pair(seed, train):
    import stereospike_b200 as sb
    from oracle import ann_ref
    torch.manual_seed(seed)
    oracle = ann_ref.AnalogUNet()
    ann_ref.randomize_batchnorm(oracle, seed=seed + 1)
    net = sb.ann.StereoSpike_equivalentANN()
    assert sorted(net.state_dict()) == sorted(oracle.state_dict())
    net.load_state_dict(oracle.state_dict())
    net = net.cuda()
    oracle.train(train)
    net.train(train)
    return oracle, net


def test_analog_model_eval_matches_oracle_and_golden(golden_dir):
    import stereospike_b200 as sb
    from oracle import make_golden as mg, ref_model as rm
    gold = np.load(os.path.join(golden_dir, 'ann_sigmoid.npz'))
    oracle, net = _pair(mg.ANN_SEED, train=False)
    x = torch.from_numpy(gold['x'].astype(np.float32))[:1]
    label = rm.synthetic_label(2, seed=mg.ANN_SEED + 200)[:1]
    with torch.no_grad():
        want = oracle(x)
        sb.functional.reset_net(net)
        got = net(x.cuda())
        again = net(x.cuda())                       # the I-neuron pool is stateful until reset, like upstream
    scale = max(float(w.abs().max()) for w in want)
    for g, w in zip(got, want):
        assert tuple(g.shape) == tuple(w.shape) == (1, 1, 260, 346)
        assert float((g.cpu() - w).abs().max()) <= TOL_DEPTH * scale, float((g.cpu() - w).abs().max()) / scale
    torch.testing.assert_close(again[3].cpu(), want[0] + want[3], rtol=0, atol=2 * TOL_DEPTH * scale)
    mde = float(rm.mean_depth_error(got[0].cpu(), label))
    assert abs(mde - float(rm.mean_depth_error(want[0], label))) <= TOL_MDE
    if mg.weights_match_arrays(mg.weight_checksum(oracle), gold['weight_checksum']):
        # the fixture comes from the reference's own ANN_models.py
        assert abs(mde - float(gold['eval_mde'])) <= TOL_MDE, (mde, float(gold['eval_mde']))
        sub = got[0][0, 0, ::4, ::4].cpu().numpy()
        assert float(np.abs(sub - gold['eval_depth1_sub']).max()) <= TOL_DEPTH * scale


def test_analog_model_training_step_matches_oracle():
    """Train mode (BatchNorm on batch statistics), B = 2: loss, every parameter gradient and the updated running statistics."""
    import stereospike_b200 as sb
    from oracle import make_golden as mg, ref_model as rm
    oracle, net = _pair(77, train=True)
    x = rm.synthetic_inputs(2, 1, 4, seed=78)
    label = rm.synthetic_label(2, seed=79)
    want = oracle(x)
    loss_o = mg.simple_loss(want, label)
    loss_o.backward()
    sb.functional.reset_net(net)
    got = net(x.cuda())
    loss = mg.simple_loss(got, label.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_o.detach())) <= 1e-4 * abs(float(loss_o.detach())), (float(loss.detach()), float(loss_o.detach()))
    g_o = dict(oracle.named_parameters())
    worst_cos, worst_rel = (1.0, None), (0.0, None)
    for k, p in net.named_parameters():
        a, b = p.grad.detach().cpu().double().flatten(), g_o[k].grad.double().flatten()
        assert float(b.norm()) > 0, k
        worst_cos = min(worst_cos, (float(torch.dot(a, b) / (a.norm() * b.norm())), k))
        worst_rel = max(worst_rel, (float((a - b).norm() / b.norm()), k))
    # Gradients in front of a train-mode BatchNorm are sums of cancelling terms (the batch statistics remove the mean), so fp32
    # convolutions in another summation order move them by up to a few 1e-3 of their norm (first layer's bias: 2e-3).
    assert worst_cos[0] >= 0.9999 and worst_rel[0] <= 1e-2, (worst_cos, worst_rel)
    for (k, a), (_, b) in zip(net.named_buffers(), oracle.named_buffers()):
        torch.testing.assert_close(a.cpu().float(), b.float(), rtol=1e-4, atol=1e-6, msg=k)


def test_analog_blocks_standalone():
    """ResBlock connect functions and the bilinear upsampling block against plain torch on the same weights."""
    import torch.nn as nn
    import torch.nn.functional as F
    import stereospike_b200 as sb
    torch.manual_seed(5)
    x = torch.randn(2, 16, 9, 11)
    for cf in ('ADD', 'MUL', 'NMUL'):
        blk = sb.ann.ResBlock(16, connect_function=cf, bias=True, activation_function=nn.Tanh()).eval()
        with torch.no_grad():
            y = blk.conv2[2](torch.tanh(blk.conv2[0](blk.conv1[2](torch.tanh(blk.conv1[0](x))))))
            want = y + x if cf == 'ADD' else (y * x if cf == 'MUL' else x * (1. - y))
            got = blk.cuda()(x.cuda()).cpu()
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    up = sb.ann.BilinConvUpsampling(16, 8, 3, (20, 24), bias=True)
    with torch.no_grad():
        want = up.up[1](F.interpolate(x, size=(22, 26), mode='bilinear', align_corners=True))
        got = up.cuda()(x.cuda()).cpu()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    with pytest.raises(NotImplementedError):
        sb.ann.ResBlock(16, connect_function='OR')
