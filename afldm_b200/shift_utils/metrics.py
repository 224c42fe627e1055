"""The equivariance metrics of /root/reference/afldm/shift_utils/metrics.py:5-19 (``mask_mse``, ``mask_psnr``,
``psnr``).  A few reductions over tensors that already live on the device; kept as PyTorch reductions on purpose
(the metric arithmetic must stay comparable with the reference's to printed precision)."""
import torch
import torch.nn.functional as F


def mask_mse(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor):
    per_sample = (a * mask - b * mask).square().sum((1, 2, 3)) / mask.sum((1, 2, 3))
    return per_sample.mean()


def mask_psnr(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor):
    am, bm = a * mask, b * mask
    i_max = torch.max(am.max(), bm.max()) - torch.min(am.min(), bm.min())
    return 10 * torch.log10(i_max * i_max / mask_mse(a, b, mask))


def psnr(a: torch.Tensor, b: torch.Tensor, i_max=None):
    if i_max is None:
        i_max = torch.max(a.max(), b.max()) - torch.min(a.min(), b.min())
    return 10 * torch.log10(i_max * i_max / F.mse_loss(a, b))
