import torch, torch.nn.functional as F
g = torch.Generator().manual_seed(0)
gt = torch.rand(2, 1, 26, 35, generator=g) * 10
gt[torch.rand(2, 1, 26, 35, generator=g) < 0.2] = float('nan')
for dev in ('cpu', 'cuda'):
    x = gt.to(dev)
    y = F.interpolate(x, size=(26, 35), mode='bilinear', align_corners=False)
    same = torch.equal(torch.nan_to_num(x, nan=-1.0), torch.nan_to_num(y, nan=-1.0))
    print(dev, 'nan in', int(torch.isnan(x).sum()), 'nan out', int(torch.isnan(y).sum()), 'identical', same)
