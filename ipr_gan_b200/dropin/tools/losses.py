"""Loss factories with the reference's surface: ``l1/mse/ms_ssim/ssim(normalized=False)`` return a
callable ``(x, y) -> 0-dim tensor`` differentiable in ``x`` (tools/loss.py:8-20, 72-85).

``ssim`` is the hot one (watermark reconstruction loss, models/wrappers.py:40): forward AND backward
run in ONE fused sm_100a kernel pass (csrc/ssim.cu); autograd only scales the stored gradient.
``l1`` / ``mse`` / ``ms_ssim`` are used by no protected config and stay as PyTorch ops on the caller's device.
"""
import torch
from torch.nn import L1Loss, MSELoss

from ipr_gan_b200 import ops

__all__ = ["l1", "mse", "ms_ssim", "ssim"]


class _FusedSSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, normalized):
        need = x.requires_grad
        loss, dx = ops.ssim_loss_fwd_bwd(x.detach(), y.detach(), normalized, 1.0, need_grad=need)
        ctx.save_for_backward(dx) if need else None
        ctx.need = need
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.need:
            return None, None, None
        (dx,) = ctx.saved_tensors
        return dx * grad_out, None, None


class _Denorm(object):
    """(x+1)/2 on both arguments before a PyTorch loss (tools/loss.py:15-20)."""

    def __init__(self, fn, normalized):
        self.fn = fn
        self.denorm = normalized

    def __call__(self, x, y):
        if self.denorm:
            x = (x + 1.0) / 2.0
            y = (y + 1.0) / 2.0
        return self.fn(x, y)


class _SSIMLoss(object):
    def __init__(self, normalized):
        self.denorm = normalized

    def __call__(self, x, y):
        if y.requires_grad:
            raise ops.IprError("ssim loss: the target must be detached (models/wrappers.py:49-51)")
        return _FusedSSIMLoss.apply(x, y, self.denorm)


def l1(normalized=False):
    return _Denorm(L1Loss(), normalized)


def mse(normalized=False):
    return _Denorm(MSELoss(), normalized)


def ssim(normalized=False):
    return _SSIMLoss(normalized)


def ms_ssim(normalized=False):
    # selected by no shipped config: outside the accelerated path, plain PyTorch composition
    import pytorch_msssim
    fn = pytorch_msssim.MS_SSIM(data_range=1)
    return _Denorm(lambda x, y: 1 - fn(x, y), normalized)
