"""The equivariance metrics of /root/reference/afldm/shift_utils/metrics.py:5-19 (``mask_mse``, ``mask_psnr``,
``psnr``).  A few reductions over tensors that already live on the device; kept as PyTorch reductions on purpose
(the metric arithmetic must stay comparable with the reference's to printed precision, so the
order of the operations - range squared, divided by the error, log10, times ten - is the reference's)."""
import torch


def _value_range(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Spread of the values of both tensors taken together: the peak signal of the PSNR (:14, :18)."""
    hi = torch.max(x.max(), y.max())
    lo = torch.min(x.min(), y.min())
    return hi - lo


def _decibel(peak, err) -> torch.Tensor:
    return 10 * torch.log10(peak * peak / err)


def mask_mse(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor):
    """Squared error summed over the valid pixels of each sample, divided by their number, averaged over the batch (:5-8)."""
    sq = (a * mask - b * mask).square()
    return (sq.sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3))).mean()


def mask_psnr(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor):
    """PSNR over the valid pixels; the peak is the value range of the MASKED tensors (:11-15)."""
    return _decibel(_value_range(a * mask, b * mask), mask_mse(a, b, mask))


def psnr(a: torch.Tensor, b: torch.Tensor, i_max=None):
    """Plain PSNR; ``i_max`` defaults to the joint value range (:17-20)."""
    peak = _value_range(a, b) if i_max is None else i_max
    return _decibel(peak, torch.nn.functional.mse_loss(a, b))
