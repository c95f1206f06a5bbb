"""White-box signature: ``SignLossModel(model, config)`` with ``forward(model)`` and
``compute_ber(model)`` (tools/sign_model.py:26-60), one kernel launch each instead of ~10 tiny
PyTorch kernels per normalisation layer.

Buffers are registered under ``name.replace('.', '_')`` in ``named_modules()`` order, exactly as
the reference does, because they are part of the checkpoint ('sign' entry).
"""
import random

import torch
import torch.nn as nn

from ipr_gan_b200 import ops


class BitGenerator:
    """Cyclic MSB-first bit stream of ``string + TAB`` (tools/sign_model.py:6-24); random bits when
    no string is given."""

    def __init__(self, string=None):
        self.random = string is None
        if string:
            assert isinstance(string, str)
            stream = []
            for ch in string + "\t":
                stream.extend(int(b) for b in format(ord(ch), "08b"))
            self.string = stream
        self.index = 0

    def __next__(self):
        if self.random:
            return random.randint(0, 1)
        bit = self.string[self.index % len(self.string)]
        self.index += 1
        return bit

    def get(self, n):
        return [next(self) for _ in range(n)]


def _signed_layers(model):
    return [(name.replace(".", "_"), m) for name, m in model.named_modules()
            if isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d))]


class _SignLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gamma_0, n, *tensors):
        gammas = [t.detach() for t in tensors[:n]]
        signs = list(tensors[n:])
        loss, grads = ops.sign_loss_fwd_bwd(gammas, signs, gamma_0)
        ctx.save_for_backward(*grads)
        ctx.n = n
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        grads = [g * grad_out for g in ctx.saved_tensors]
        return (None, None) + tuple(grads) + (None,) * ctx.n


class SignLossModel(nn.Module):
    def __init__(self, model, config, **kwargs):
        super().__init__()
        self.gamma_0 = config.gamma_0
        self.bit_gen = BitGenerator(config.string)
        self._create_signs(model)

    def _create_signs(self, model):
        for safe, m in _signed_layers(model):
            bits = torch.tensor(self.bit_gen.get(m.weight.size(0)), dtype=torch.float32)
            sign = bits * 2 - 1
            with torch.no_grad():
                m.weight.abs_().mul_(sign.to(m.weight.device))
            self.register_buffer(safe, sign)

    def _collect(self, model):
        gammas, signs = [], []
        for safe, m in _signed_layers(model):
            gammas.append(m.weight)
            signs.append(getattr(self, safe))
        return gammas, signs

    def forward(self, model):
        gammas, signs = self._collect(model)
        return _SignLoss.apply(self.gamma_0, len(gammas), *gammas, *signs)

    def value_into(self, model, slot, scale=1.0):
        """slot[0] <- scale * sign loss, no gradient output (fused path: d/dgamma is added by the normalisation
        layers' backward kernels, models/protect.py)."""
        gammas, signs = self._collect(model)
        with torch.no_grad():
            ops.sign_loss_fwd_bwd([g.detach() for g in gammas], signs, self.gamma_0, need_grad=False, loss_out=slot,
                                  loss_scale=scale)

    def compute_ber(self, model):
        gammas, signs = self._collect(model)
        with torch.no_grad():
            counts = ops.sign_ber_counts([g.detach() for g in gammas], signs)
            return counts[0].float() / counts[1].float()

    def compute_ber_counts(self, model):
        """(wrong, total) as Python ints -- the bit-exact quantity behind compute_ber."""
        gammas, signs = self._collect(model)
        with torch.no_grad():
            c = ops.sign_ber_counts([g.detach() for g in gammas], signs).tolist()
        return int(c[0]), int(c[1])
