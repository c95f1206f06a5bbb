"""oracle/shims/pytorch_msssim -- TEST INFRASTRUCTURE ONLY.

CPU/PyTorch restatement of the two entry points of the third-party package
``pytorch-msssim==0.2.1`` (pinned at /root/reference/requirements.txt:6, upstream
VainF/pytorch-msssim) that the reference calls on the hot path:

* ``SSIM(data_range=1)``                       -- /root/reference/tools/loss.py:3,83
* ``ssim(X, Y, data_range=1, size_average=False)`` -- /root/reference/experiments/image_generation.py:6,211

The package is absent from this image and from /root/reference, so its published
algorithm is restated here (PARITY UNPINNED -- the reference ships no SSIM test vectors):

* window: 11 taps, sigma 1.5, ``g = exp(-(i-5)^2 / (2 sigma^2)); g /= g.sum()`` in fp32;
* filtering: depthwise *valid* (un-padded) correlation, separable, H pass first then W,
  a spatial dim shorter than the window is skipped;
* statistics: mu_x, mu_y, G*x^2 - mu_x^2, G*y^2 - mu_y^2, G*xy - mu_x mu_y;
* ``cs = (2 s_xy + C2) / (s_x + s_y + C2)``, ``ssim = (2 mu_x mu_y + C1)/(mu_x^2 + mu_y^2 + C1) * cs``
  with C1 = (0.01 L)^2, C2 = (0.03 L)^2;
* reduction: mean over the valid map per (n, c), then mean over c (per sample) or over (n, c).

It is put on ``sys.path`` only by ``oracle/ref_bridge.py`` so that the reference's own
``tools`` / ``experiments`` packages import on CPU.
"""
import warnings

import torch
import torch.nn.functional as F

__version__ = "0.2.1-oracle-restatement"


def gauss_taps(size=11, sigma=1.5):
    pos = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(pos ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _blur(t, taps):
    """Depthwise valid correlation with the 1-D taps along H, then along W."""
    ch = t.shape[1]
    n = taps.numel()
    k = taps.to(t.device, t.dtype)
    if t.shape[2] >= n:
        t = F.conv2d(t, k.view(1, 1, n, 1).repeat(ch, 1, 1, 1), groups=ch)
    else:
        warnings.warn("ssim: H shorter than the window, H pass skipped")
    if t.shape[3] >= n:
        t = F.conv2d(t, k.view(1, 1, 1, n).repeat(ch, 1, 1, 1), groups=ch)
    else:
        warnings.warn("ssim: W shorter than the window, W pass skipped")
    return t


def _ssim_and_cs(x, y, data_range, taps, K):
    c1 = (K[0] * data_range) ** 2
    c2 = (K[1] * data_range) ** 2
    mu_x = _blur(x, taps)
    mu_y = _blur(y, taps)
    mu_xx = mu_x.pow(2)
    mu_yy = mu_y.pow(2)
    mu_xy = mu_x * mu_y
    var_x = _blur(x * x, taps) - mu_xx
    var_y = _blur(y * y, taps) - mu_yy
    cov = _blur(x * y, taps) - mu_xy
    cs_map = (2 * cov + c2) / (var_x + var_y + c2)
    s_map = ((2 * mu_xy + c1) / (mu_xx + mu_yy + c1)) * cs_map
    return torch.flatten(s_map, 2).mean(-1), torch.flatten(cs_map, 2).mean(-1)


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None,
         K=(0.01, 0.03), nonnegative_ssim=False):
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    if X.dim() != 4:
        raise ValueError("this restatement covers (N, C, H, W) inputs only")
    if win_size % 2 != 1:
        raise ValueError("Window size should be odd.")
    taps = gauss_taps(win_size, win_sigma) if win is None else win.flatten()
    per_ch, _ = _ssim_and_cs(X, Y, data_range, taps, K)
    if nonnegative_ssim:
        per_ch = torch.relu(per_ch)
    return per_ch.mean() if size_average else per_ch.mean(1)


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None,
            weights=None, K=(0.01, 0.03)):
    if X.shape != Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    taps = gauss_taps(win_size, win_sigma) if win is None else win.flatten()
    if weights is None:
        weights = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
    w = torch.tensor(weights, dtype=X.dtype, device=X.device)
    smaller = min(X.shape[-2:])
    assert smaller > (win_size - 1) * (2 ** 4), "image too small for 5-level MS-SSIM"
    terms = []
    for lvl in range(w.numel()):
        s, cs = _ssim_and_cs(X, Y, data_range, taps, K)
        if lvl < w.numel() - 1:
            terms.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X = F.avg_pool2d(X, 2, padding=pad)
            Y = F.avg_pool2d(Y, 2, padding=pad)
    terms.append(torch.relu(s))
    stack = torch.stack(terms, dim=0)
    val = torch.prod(stack ** w.view(-1, 1, 1), dim=0)
    return val.mean() if size_average else val.mean(1)


class SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3,
                 spatial_dims=2, K=(0.01, 0.03), nonnegative_ssim=False):
        super().__init__()
        self.win_size = win_size
        self.win = gauss_taps(win_size, win_sigma)
        self.size_average = size_average
        self.data_range = data_range
        self.K = K
        self.nonnegative_ssim = nonnegative_ssim

    def forward(self, X, Y):
        return ssim(X, Y, data_range=self.data_range, size_average=self.size_average, win=self.win,
                    K=self.K, nonnegative_ssim=self.nonnegative_ssim)


class MS_SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3,
                 spatial_dims=2, weights=None, K=(0.01, 0.03)):
        super().__init__()
        self.win_size = win_size
        self.win = gauss_taps(win_size, win_sigma)
        self.size_average = size_average
        self.data_range = data_range
        self.weights = weights
        self.K = K

    def forward(self, X, Y):
        return ms_ssim(X, Y, data_range=self.data_range, size_average=self.size_average, win=self.win,
                       weights=self.weights, K=self.K)
