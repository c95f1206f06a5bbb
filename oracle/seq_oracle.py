"""oracle/seq_oracle.py -- TEST INFRASTRUCTURE ONLY (checker; never the thing measured or shipped).

PyTorch restatement of how the SRGAN / CycleGAN network families are evaluated, for parity tests of
``ipr_gan_b200.seqnet``:

* ``forward_fp32(net, x)``: the module tree with plain PyTorch operators in fp32 -- exactly what the reference's
  ``nn.Sequential`` subclasses compute (networks/sr_resnet.py:3-44, discriminator_96.py:3-35,
  resnet_generator.py:3-59, conv_discriminator.py:3-21);
* ``forward_sim_bf16(net, x)``: the SAME PyTorch operators with values rounded to bf16 at exactly the points where
  the sm_100a engine stores bf16 (packed weights, the convolution output, every block output, and the gradients at
  the same points) and fp32 everywhere else.  Against this model the engine must agree to accumulation-order noise;
  in particular the ReLU / LeakyReLU / PReLU on/off patterns coincide, because both sides take the sign of the same
  bf16-rounded pre-activation -- which is what lets whole-network gradients be held to the bf16 budget (2e-2).

The block structure is read from ``ipr_gan_b200.seqnet.lower`` (lists of conv / norm / activation modules); all
arithmetic here is torch's.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwd(torch.autograd.Function):          # weights: rounded copy, straight-through gradient
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def forward_fp32(net, x):
    y = nn.Sequential.forward(net, x)
    return y.squeeze() if type(net).__name__ == "Discriminator96" else y


def _gate(y, mask, slope):
    """Activation with a PRESCRIBED on/off pattern: y where mask else slope * y, forward and backward."""
    return y * torch.where(mask, torch.ones((), dtype=y.dtype), slope * torch.ones((), dtype=y.dtype))


def forward_sim_bf16(net, x, blocks, capture=None, masks=None):
    """``blocks`` = ipr_gan_b200.seqnet.lower(net).  x: NCHW fp32.  ``capture``: optional list receiving every block's
    (rounded convolution output, block output) in NCHW.  ``masks``: optional per-block boolean NCHW tensors (None for
    blocks without a piecewise-linear activation) prescribing the ReLU / LeakyReLU / PReLU on/off pattern -- the
    engine's own.  Two bf16 evaluations of a deep network differ by about one bf16 ulp per layer (each side rounds a
    slightly different fp32 value), which moves ~1 % of the pre-activations across zero; every such element changes
    its gradient by 100 %.  With the pattern prescribed, what remains is arithmetic."""
    r, wq = _RoundBoth.apply, _RoundFwd.apply
    tensors = [_RoundFwd.apply(x)]
    out = None
    for b in blocks:
        t = tensors[-1]
        conv = b.conv
        w = wq(conv.weight)
        if isinstance(conv, nn.ConvTranspose2d):
            y = F.conv_transpose2d(t, w, None, conv.stride, conv.padding, conv.output_padding)
        else:
            if b.reflect:
                t = F.pad(t, (b.pad,) * 4, mode="reflect")
                y = F.conv2d(t, w, None, conv.stride, 0)
            else:
                y = F.conv2d(t, w, None, conv.stride, conv.padding)
        if b.final:
            if conv.bias is not None:
                y = y + conv.bias.view(1, -1, 1, 1)
            out = torch.tanh(y) if b.act == 4 else y
            break
        if conv.bias is not None:
            y = y + conv.bias.view(1, -1, 1, 1)
        y = r(y)                                                   # the engine stores the convolution output in bf16
        y_conv = y
        nm = b.norm
        if isinstance(nm, nn.BatchNorm2d):
            use_batch = nm.training or (nm.running_mean is None and nm.running_var is None)
            upd = nm.training and nm.track_running_stats
            y = F.batch_norm(y, nm.running_mean if (upd or not use_batch) else None,
                             nm.running_var if (upd or not use_batch) else None, nm.weight, nm.bias, use_batch,
                             nm.momentum, nm.eps)
            if upd:
                nm.num_batches_tracked += 1
        elif isinstance(nm, nn.InstanceNorm2d):
            y = F.instance_norm(y, None, None, nm.weight, nm.bias, True, 0.1, nm.eps)
        m = masks[len(tensors) - 1] if masks is not None else None
        if b.act == 1:
            y = F.relu(y) if m is None else _gate(y, m, 0.0)
        elif b.act == 2:
            y = F.leaky_relu(y, b.slope) if m is None else _gate(y, m, b.slope)
        elif b.act == 3:
            y = F.prelu(y, b.prelu.weight) if m is None else _gate(y, m, b.prelu.weight)
        elif b.act == 4:
            y = torch.tanh(y)
        if b.residual is not None:
            y = y + tensors[b.residual]
        y = r(y)
        if b.shuffle:
            y = F.pixel_shuffle(y, 2)
        if capture is not None:
            capture.append((y_conv.detach(), y.detach()))
        tensors.append(y)
    return out.squeeze() if type(net).__name__ == "Discriminator96" else out
