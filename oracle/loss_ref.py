"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement of the reference's training loss and depth metric: network/loss.py:7-135 (ScaleInvariant_Loss,
Multiscale_ScaleInvariant_Loss, GradientMatching_Loss, MultiScale_GradientMatching_Loss, SpikePenalization_Loss,
Total_Loss) and network/metrics.py:83-95 (MeanDepthError), in plain PyTorch and device-agnostic (the reference moves its
Sobel filters to CUDA whenever a GPU exists, loss.py:60-65, whatever the device of its inputs).  Pinned by executing the
reference's own files on the same inputs: tests/test_oracle_loss.py.
"""
import torch
import torch.nn.functional as F


def scale_invariant_loss(predicted, groundtruth):                      # loss.py:7-24
    mask = ~torch.isnan(groundtruth)
    n = torch.count_nonzero(mask)
    res = predicted - groundtruth
    res[mask == False] = 0                                             # noqa: E712  (as the reference writes it)
    mse = 1 / n * torch.sum(torch.pow(res[mask], 2))
    quad = 1 / (n ** 2) * torch.pow(torch.sum(res[mask]), 2)
    return mse - quad


def gradient_matching_loss(predicted, groundtruth):                    # loss.py:44-76
    mask = ~torch.isnan(groundtruth)
    n = torch.count_nonzero(mask)
    res = predicted - groundtruth
    res[mask == False] = 0                                             # noqa: E712
    sobel_x = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], dtype=res.dtype, device=res.device).view(1, 1, 3, 3)
    sobel_y = torch.tensor([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=res.dtype, device=res.device).view(1, 1, 3, 3)
    gx = F.conv2d(res, sobel_x, stride=1, padding=1)
    gy = F.conv2d(res, sobel_y, stride=1, padding=1)
    gx = gx * mask
    gy = gy * mask
    return 1 / n * torch.sum(torch.abs(gx[mask]) + torch.abs(gy[mask]))


def _multiscale(fn, predicted, groundtruth, factors):                  # loss.py:27-41, 79-93
    total = 0.0
    for factor, m in zip(factors, predicted):
        gt = F.interpolate(groundtruth, size=(m.shape[-2], m.shape[-1]), mode='bilinear', align_corners=False)
        total = total + factor * fn(m, gt)
    return total


def spike_penalization_loss(spike_tensors):                            # loss.py:96-107
    total = 0.0
    for s in spike_tensors:
        total = total + 1 / (2 * s.numel()) * torch.sum(torch.pow(s, 2))
    return total


def total_loss(predicted, groundtruth, alpha=0.5, scale_weights=(1., 1., 1., 1.), spikes=None, beta=1.):   # loss.py:110-135
    loss = _multiscale(scale_invariant_loss, predicted, groundtruth, scale_weights) + \
        alpha * _multiscale(gradient_matching_loss, predicted, groundtruth, scale_weights)
    if spikes is not None:
        loss = loss + beta * spike_penalization_loss(spikes)
    return loss


def mean_depth_error(predicted, groundtruth):                          # metrics.py:83-95
    mask = ~torch.isnan(groundtruth)
    n = torch.count_nonzero(mask)
    res = predicted - groundtruth
    res[mask == False] = 0                                             # noqa: E712
    return torch.sum(torch.abs(res[mask])) / n
