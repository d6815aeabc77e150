"""Drop-in for the reference's ``network/loss.py`` and ``network/metrics.py:MeanDepthError`` on the fused CUDA passes
``ss_loss_fwd`` / ``ss_loss_bwd`` (include/stereospike_b200.h): same function / class names, arguments and results.

The reference evaluates every scale with ~10 PyTorch kernels and boolean-mask gathers (loss.py:17-24, 53-76) and rescales
the ground truth with ``F.interpolate`` (loss.py:38) -- an identity here, because all four depth maps are at full
resolution (SNN_models.py:133-148).  Predictions whose size differs from the ground truth are still supported: the ground
truth is then interpolated exactly as in the reference and that scale is evaluated by its own call.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .ops import _ptr, _require_cuda, _stream


def _maps_bhw(t, name):
    _require_cuda(t, name)
    if t.dim() == 4:
        assert t.shape[1] == 1, f'{name}: expected [N, 1, H, W]'
        t = t[:, 0]
    assert t.dim() == 3, f'{name}: expected [N, 1, H, W] or [N, H, W]'
    return t.contiguous().float()


def _fwd(preds, gt):
    """preds: list of <= 4 fp32 [B,H,W]; gt fp32 [B,H,W].  Returns (sums fp64 [n,5], signs u8 [n,B*H*W])."""
    B, H, W = gt.shape
    n = len(preds)
    sums = torch.zeros((n, 5), dtype=torch.float64, device=gt.device)
    signs = torch.empty((n, B * H * W), dtype=torch.uint8, device=gt.device)
    arr = (ctypes.c_void_p * 4)(*[p.data_ptr() for p in preds] + [0] * (4 - n))
    _lib.check(_lib.lib().ss_loss_fwd(n, B, H, W, arr, _ptr(gt), _ptr(sums), _ptr(signs), _stream()), 'ss_loss_fwd')
    return sums, signs


class _FusedLoss(torch.autograd.Function):
    """sum_k w_si[k] * SI_k + w_gm[k] * GM_k over predictions that share the ground truth's size."""

    @staticmethod
    def forward(ctx, gt, w_si, w_gm, *preds):
        ps = [_maps_bhw(p, 'predicted') for p in preds]
        sums, signs = _fwd(ps, gt)
        n, s1, s2, g = sums[:, 0], sums[:, 1], sums[:, 2], sums[:, 3]
        si = s2 / n - (s1 / n) ** 2
        gm = g / n
        wsi = torch.tensor(w_si, dtype=torch.float64, device=gt.device)
        wgm = torch.tensor(w_gm, dtype=torch.float64, device=gt.device)
        ctx.save_for_backward(gt, sums, signs, *ps)
        ctx.w_si, ctx.w_gm, ctx.shapes = w_si, w_gm, [p.shape for p in preds]
        mde = (sums[:, 4] / n).float()                  # MeanDepthError of every scale, from the same pass (metrics.py:83-95)
        ctx.mark_non_differentiable(mde)
        return (wsi * si + wgm * gm).sum().float(), mde

    @staticmethod
    def backward(ctx, g_out, _g_mde=None):
        gt, sums, signs, *ps = ctx.saved_tensors
        n = len(ps)
        B, H, W = gt.shape
        grads = [torch.empty((B, H, W), dtype=torch.float32, device=gt.device) for _ in range(n)]
        g = g_out.float().reshape(1)
        c_si = torch.tensor(ctx.w_si, dtype=torch.float32, device=gt.device) * g
        c_gm = torch.tensor(ctx.w_gm, dtype=torch.float32, device=gt.device) * g
        parr = (ctypes.c_void_p * 4)(*[p.data_ptr() for p in ps] + [0] * (4 - n))
        garr = (ctypes.c_void_p * 4)(*[x.data_ptr() for x in grads] + [0] * (4 - n))
        _lib.check(_lib.lib().ss_loss_bwd(n, B, H, W, parr, _ptr(gt), _ptr(sums), _ptr(signs), _ptr(c_si), _ptr(c_gm), garr,
                                          _stream()), 'ss_loss_bwd')
        return (None, None, None) + tuple(x.view(s) for x, s in zip(grads, ctx.shapes))


def _terms(predicted, groundtruth, w_si, w_gm):
    """Shared driver: groups the predictions by size (the ground truth is interpolated per distinct size, loss.py:38)."""
    gt_full = groundtruth
    _require_cuda(gt_full, 'groundtruth')
    total = None
    mdes = {}
    by_size = {}
    for k, m in enumerate(predicted):
        by_size.setdefault((m.shape[-2], m.shape[-1]), []).append(k)
    for size, ks in by_size.items():
        gt = gt_full if tuple(gt_full.shape[-2:]) == size else F.interpolate(gt_full, size=size, mode='bilinear', align_corners=False)
        gt = _maps_bhw(gt, 'groundtruth')
        for i in range(0, len(ks), 4):
            chunk = ks[i:i + 4]
            val, mde = _FusedLoss.apply(gt, tuple(float(w_si[k]) for k in chunk), tuple(float(w_gm[k]) for k in chunk),
                                        *[predicted[k] for k in chunk])
            for j, k in enumerate(chunk):
                mdes[k] = mde[j]
            total = val if total is None else total + val
    return total, mdes


def ScaleInvariant_Loss(predicted, groundtruth):
    """loss.py:7-24."""
    return _terms([predicted], groundtruth, [1.0], [0.0])[0]


def GradientMatching_Loss(predicted, groundtruth):
    """loss.py:44-76."""
    return _terms([predicted], groundtruth, [0.0], [1.0])[0]


def Multiscale_ScaleInvariant_Loss(predicted, groundtruth, factors=(1., 1., 1., 1.)):
    """loss.py:27-41."""
    ps = list(predicted)[:len(factors)]
    return _terms(ps, groundtruth, list(factors)[:len(ps)], [0.0] * len(ps))[0]


def MultiScale_GradientMatching_Loss(predicted, groundtruth, factors=(1., 1., 1., 1.)):
    """loss.py:79-93."""
    ps = list(predicted)[:len(factors)]
    f = list(factors)[:len(ps)]
    return _terms(ps, groundtruth, [0.0] * len(ps), f)[0]


def SpikePenalization_Loss(intermediary_spike_tensors):
    """loss.py:96-107.  The spike maps returned by ``forward`` / ``forward_seq(spikes_fp32=True)`` are outputs of the
    network's autograd node, so this term back-propagates through the surrogates like in the reference."""
    loss = 0.0
    for s in intermediary_spike_tensors:
        s = s.float()
        loss = loss + 1 / (2 * s.numel()) * torch.sum(torch.pow(s, 2))
    return loss


class Total_Loss(nn.Module):
    """loss.py:110-135: multi-scale scale-invariant loss + alpha * multi-scale gradient-matching loss
    (+ beta * spike penalisation).  ``last_mde`` holds the MeanDepthError of ``predicted[0]`` from the same pass."""

    def __init__(self, alpha=0.5, scale_weights=(1., 1., 1., 1.), penalize_spikes=False, beta=1.):
        super().__init__()
        self.alpha = alpha
        self.scale_weights = scale_weights
        self.penalize_spikes = penalize_spikes
        self.beta = beta
        self.last_mde = None

    def forward(self, predicted, groundtruth, intermediary_spike_tensors=None):
        ps = list(predicted)[:len(self.scale_weights)]
        w = [float(f) for f in list(self.scale_weights)[:len(ps)]]
        loss, mdes = _terms(ps, groundtruth, w, [self.alpha * f for f in w])
        self.last_mde = mdes[0].detach()                # device scalar: no host sync unless the caller reads it
        if self.penalize_spikes:
            fused = getattr(intermediary_spike_tensors, 'penalty', None)      # models.SpikeList: computed in the block epilogues
            loss = loss + self.beta * (fused if fused is not None else SpikePenalization_Loss(intermediary_spike_tensors))
        return loss


def MeanDepthError(predicted, groundtruth):
    """metrics.py:83-95: mean |predicted - groundtruth| over the non-NaN ground-truth pixels (no gradient)."""
    with torch.no_grad():
        gt = _maps_bhw(groundtruth, 'groundtruth')
        p = _maps_bhw(predicted.detach(), 'predicted')
        sums, _ = _fwd([p], gt)
        return (sums[0, 4] / sums[0, 0]).float()
