"""Drop-in for the two ``pytorch_msssim`` entry points the reference imports
(tools/loss.py:3, experiments/image_generation.py:6), served by the fused sm_100a SSIM kernels.

Only what the reference calls is provided: ``ssim(X, Y, data_range=1, size_average=...)`` with the
default 11-tap sigma-1.5 window, and ``SSIM(data_range=1)``.  Host tensors are staged to the
current CUDA device, computed there and the result is returned on the inputs' device.

``ms_ssim`` / ``MS_SSIM`` (tools/loss.py:78-80; selected by no shipped config) are outside the
accelerated path: a plain PyTorch composition (depthwise Gaussian filters + 2x average pooling, five scales) on the
inputs' own device, kept so that ``loss_fn: ms_ssim`` keeps working.
"""
import torch

from ipr_gan_b200 import ops

__all__ = ["ssim", "SSIM", "ms_ssim", "MS_SSIM"]


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


# ------------------------------------------------------------------------------------------------ multi-scale
_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _window(size, sigma, like):
    pos = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(pos ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).to(like.device, like.dtype)


def _smooth(t, win):
    c, n = t.shape[1], win.numel()
    if t.shape[2] >= n:
        t = torch.nn.functional.conv2d(t, win.view(1, 1, n, 1).expand(c, 1, n, 1), groups=c)
    if t.shape[3] >= n:
        t = torch.nn.functional.conv2d(t, win.view(1, 1, 1, n).expand(c, 1, 1, n), groups=c)
    return t


def _scale_terms(X, Y, data_range, win, K):
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mx, my = _smooth(X, win), _smooth(Y, win)
    vx = _smooth(X * X, win) - mx * mx
    vy = _smooth(Y * Y, win) - my * my
    vxy = _smooth(X * Y, win) - mx * my
    contrast = (2 * vxy + c2) / (vx + vy + c2)
    full = (2 * mx * my + c1) / (mx * mx + my * my + c1) * contrast
    return full.flatten(2).mean(-1), contrast.flatten(2).mean(-1)


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, weights=None,
            K=(0.01, 0.03)):
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    if min(X.shape[-2:]) <= (win_size - 1) * 2 ** 4:
        raise AssertionError("Image size should be larger than %d due to the 4 downsamplings in ms-ssim"
                             % ((win_size - 1) * 2 ** 4))
    win = _window(win_size, win_sigma, X) if win is None else win.flatten().to(X.device, X.dtype)
    w = torch.tensor(list(weights or _MS_WEIGHTS), device=X.device, dtype=X.dtype)
    levels = []
    for lvl in range(w.numel()):
        full, contrast = _scale_terms(X, Y, data_range, win, K)
        last = lvl == w.numel() - 1
        levels.append(torch.relu(full if last else contrast))
        if not last:
            pad = [s % 2 for s in X.shape[2:]]
            X = torch.nn.functional.avg_pool2d(X, 2, padding=pad)
            Y = torch.nn.functional.avg_pool2d(Y, 2, padding=pad)
    val = torch.prod(torch.stack(levels) ** w.view(-1, 1, 1), dim=0)
    return val.mean() if size_average else val.mean(1)


class MS_SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 weights=None, K=(0.01, 0.03)):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma,
                       weights=weights, K=K)

    def forward(self, X, Y):
        return ms_ssim(X, Y, **self.kw)
