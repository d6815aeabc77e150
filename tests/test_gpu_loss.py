"""Fused loss / metric passes (ss_loss_fwd, ss_loss_bwd through stereospike_b200.loss) against the oracle
(oracle/loss_ref.py, pinned by the reference's own loss.py in tests/test_oracle_loss.py).  Floating point: the two sides
sum ~1e5 fp32 terms in different orders (ours in fp64), tolerance 2e-5 relative on values, 1e-5 * max|grad| on gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(seed, B, H, W, nan_frac=0.2):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, H, W, generator=g) * 10
    gt[torch.rand(B, 1, H, W, generator=g) < nan_frac] = float('nan')
    preds = [torch.rand(B, 1, H, W, generator=g) * 10 for _ in range(4)]
    return preds, gt


@pytest.mark.parametrize('B,H,W,weights,alpha', [(2, 26, 35, (1., 1., 1., 1.), 0.5), (3, 260, 346, (1., 0.5, 2., 0.25), 0.5),
                                                 (1, 7, 5, (1., 1.), 2.0)])
def test_total_loss_value_and_gradients(B, H, W, weights, alpha):
    from oracle import loss_ref
    from stereospike_b200 import loss as sl
    preds, gt = _inputs(5, B, H, W)
    preds = preds[:len(weights)]
    pa = [p.clone().requires_grad_(True) for p in preds]
    la = loss_ref.total_loss(pa, gt, alpha=alpha, scale_weights=weights)
    (la * 1.7).backward()
    pb = [p.clone().cuda().requires_grad_(True) for p in preds]
    lb = sl.Total_Loss(alpha=alpha, scale_weights=weights)(pb, gt.cuda())
    (lb * 1.7).backward()
    torch.cuda.synchronize()
    assert abs(float(la.detach()) - float(lb.detach())) <= 2e-5 * abs(float(la.detach())), (float(la.detach()), float(lb.detach()))
    for a, b in zip(pa, pb):
        scale = float(a.grad.abs().max())
        assert float((a.grad - b.grad.cpu()).abs().max()) <= 1e-5 * scale + 1e-12
        assert torch.equal(b.grad.cpu()[torch.isnan(gt)], torch.zeros_like(b.grad.cpu()[torch.isnan(gt)]))


def test_single_terms_metric_and_edge_cases():
    from oracle import loss_ref
    from stereospike_b200 import loss as sl
    preds, gt = _inputs(9, 2, 33, 44)
    rel = lambda a, b: abs(float(a) - float(b)) <= 2e-5 * abs(float(a)) + 1e-9
    assert rel(loss_ref.scale_invariant_loss(preds[0], gt), sl.ScaleInvariant_Loss(preds[0].cuda(), gt.cuda()))
    assert rel(loss_ref.gradient_matching_loss(preds[1], gt), sl.GradientMatching_Loss(preds[1].cuda(), gt.cuda()))
    assert rel(loss_ref.mean_depth_error(preds[2], gt), sl.MeanDepthError(preds[2].cuda(), gt.cuda()))
    f = (1., 2., 3., 4.)
    assert rel(loss_ref._multiscale(loss_ref.scale_invariant_loss, preds, gt, f), sl.Multiscale_ScaleInvariant_Loss([p.cuda() for p in preds], gt.cuda(), f))
    assert rel(loss_ref._multiscale(loss_ref.gradient_matching_loss, preds, gt, f), sl.MultiScale_GradientMatching_Loss([p.cuda() for p in preds], gt.cuda(), f))
    # no invalid pixels at all; a prediction at another resolution (the ground truth is interpolated as in the reference)
    gt2 = torch.nan_to_num(gt, nan=1.0)
    assert rel(loss_ref.total_loss(preds, gt2), sl.Total_Loss()([p.cuda() for p in preds], gt2.cuda()))
    small = [preds[0], torch.rand(2, 1, 17, 22) * 10]
    assert rel(loss_ref.total_loss(small, gt2, scale_weights=(1., 1.)), sl.Total_Loss(scale_weights=(1., 1.))([p.cuda() for p in small], gt2.cuda()))
    # a CPU tensor is an error, not a fallback
    with pytest.raises(RuntimeError):
        sl.Total_Loss()(preds, gt)


def test_training_step_with_fused_loss_matches_oracle_loss():
    """Gradients of the network's parameters under the fused loss == under the oracle's loss evaluated on the same depth maps.
    Uses the fp32 CUDA-core backward, which is reproducible to ~1e-6 run to run; the bf16 tensor-core backward amplifies
    the atomics' summation-order noise to its own rounding level (<= 0.5 % of a tensor's largest gradient), which would
    mask what this test is about."""
    import stereospike_b200 as sb
    from oracle import loss_ref, ref_model as rm
    from stereospike_b200 import loss as sl
    torch.manual_seed(3)
    net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0).cuda()
    net.set_kernel_options(bwd_impl='simt')
    x = rm.synthetic_inputs(1, 2, 4, seed=6).cuda()
    label = rm.synthetic_label(1, seed=7).cuda()
    grads = []
    for fused in (True, False):
        net.zero_grad()
        sb.functional.reset_net(net)
        pred, _ = net.forward_seq(x)
        loss = sl.Total_Loss()(pred, label) if fused else loss_ref.total_loss(pred, label)
        loss.backward()
        grads.append([p.grad.clone() for p in net.parameters()])
    for a, b in zip(*grads):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-10


def test_spike_penalisation_back_propagates_like_the_oracle():
    """Total_Loss(penalize_spikes=True) (loss.py:96-135): the penalty on the returned spike maps reaches the weights through
    the surrogate gradients.  Parameter gradients against autograd through the oracle (cosine >= 0.999, as for the depth
    loss), and the penalty must actually change them."""
    import stereospike_b200 as sb
    from oracle import loss_ref, ref_model as rm, sj_compat as sj
    from stereospike_b200 import loss as sl
    torch.manual_seed(5)     # a configuration on which the fp32 oracle and the exact-integer forward agree spike for spike
    oracle = rm.SpikingUNet('lif', tau=3.0, multiply_factor=15.0)
    net = sb.fromZero_feedforward_multiscale_tempo_Matt_SpikeFlowNetLike(use_plif=False, tau=3.0, multiply_factor=15.0)
    net.load_state_dict(oracle.state_dict())
    net = net.cuda()
    x = rm.synthetic_inputs(1, 2, 4, seed=6)
    label = rm.synthetic_label(1, seed=7)
    beta = 50.0
    sj.reset_net(oracle)
    d_ref, s_ref = oracle.forward_seq(x)
    loss_ref.total_loss(d_ref, label, spikes=s_ref, beta=beta).backward()
    got = {}
    for penal in (True, False):
        net.zero_grad()
        sb.functional.reset_net(net)
        d, s = net.forward_seq(x.cuda(), spikes_fp32=True)
        # the fused penalty (counters accumulated in the block epilogues) is the reference formula on the returned maps
        assert s.penalty is not None and s.penalty.requires_grad
        torch.testing.assert_close(s.penalty.detach(), sl.SpikePenalization_Loss(list(s)).detach(), rtol=1e-6, atol=0)
        sl.Total_Loss(penalize_spikes=penal, beta=beta)(d, label.cuda(), s).backward()
        got[penal] = {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}
    ref = dict(oracle.named_parameters())
    changed = 0
    for k, g in got[True].items():
        a, b = ref[k].grad.flatten().double(), g.flatten().double()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos >= 0.999, (k, cos)
        changed += int(float((g - got[False][k]).abs().max()) > 1e-3 * float(g.abs().max()))
    assert changed >= 10, changed
