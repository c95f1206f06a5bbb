"""Drop-in for the two ``pytorch_msssim`` entry points the reference imports
(tools/loss.py:3, experiments/image_generation.py:6), served by the fused sm_100a SSIM kernels.

Only what the reference calls is provided: ``ssim(X, Y, data_range=1, size_average=...)`` with the
default 11-tap sigma-1.5 window, and ``SSIM(data_range=1)``.  Host tensors are staged to the
current CUDA device, computed there and the result is returned on the inputs' device.
"""
import torch

from ipr_gan_b200 import ops

__all__ = ["ssim", "SSIM"]


def _check(data_range, win_size, win_sigma, win, K, nonnegative_ssim):
    if data_range != 1 or win_size != 11 or win_sigma != 1.5 or win is not None or tuple(K) != (0.01, 0.03) \
            or nonnegative_ssim:
        raise NotImplementedError("only data_range=1, win 11/1.5, K=(0.01,0.03) (the reference's call) is built")


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03),
         nonnegative_ssim=False):
    _check(data_range, win_size, win_sigma, win, K, nonnegative_ssim)
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    src = X.device
    dev = src if X.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if size_average:
        from tools.losses import _FusedSSIMLoss
        return (1 - _FusedSSIMLoss.apply(X.to(dev, torch.float32), Y.detach().to(dev, torch.float32), False)).to(src)
    with torch.no_grad():
        out = ops.ssim_per_sample(X.detach().to(dev, torch.float32), Y.detach().to(dev, torch.float32))
    return out.to(src)


class SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 K=(0.01, 0.03), nonnegative_ssim=False):
        super().__init__()
        _check(data_range, win_size, win_sigma, None, K, nonnegative_ssim)
        self.data_range = data_range
        self.size_average = size_average

    def forward(self, X, Y):
        return ssim(X, Y, data_range=self.data_range, size_average=self.size_average)
