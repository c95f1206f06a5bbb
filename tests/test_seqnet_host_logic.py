"""CPU: the host-side planning of the generic convolution path (ipr_gan_b200/seqnet.py) -- block lowering, weight
packing for the forward / data-gradient GEMMs, the weight-gradient scatter tables -- together with an fp64 PyTorch
emulation of the kernels' contract (csrc/layers.cu: patch matrix, adjoint gather; one single-tap GEMM in between),
against torch's own convolutions for every layer shape of the SRGAN / CycleGAN networks
(networks/sr_resnet.py, discriminator_96.py, resnet_generator.py, conv_discriminator.py)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


def _src(v, size, up, reflect):
    """csrc/layers.cu: src_index"""
    if reflect:
        if v < 0:
            v = -v
        if v > size - 1:
            v = 2 * (size - 1) - v
        return v
    if v < 0 or v > (size - 1) * up:
        return -1
    if up == 1:
        return v
    return v // up if v % up == 0 else -1


def emulate_im2col(x, b, oh, ow):
    """csrc/layers.cu: im2col_kernel (x: NHWC, fp64)"""
    n, h, w, c = x.shape
    col = torch.zeros(n * oh * ow, b.kp, dtype=x.dtype)
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(b.k):
                for kx in range(b.k):
                    iy = _src(oy * b.stride + ky - b.pad, h, b.up, b.reflect)
                    ix = _src(ox * b.stride + kx - b.pad, w, b.up, b.reflect)
                    if iy >= 0 and ix >= 0:
                        rows = torch.arange(n) * oh * ow + oy * ow + ox
                        col[rows, (ky * b.k + kx) * c:(ky * b.k + kx + 1) * c] = x[:, iy, ix, :]
    return col


def emulate_col2im(dcol, b, shape, oh, ow):
    """csrc/layers.cu: col2im_kernel"""
    n, h, w, c = shape
    dx = torch.zeros(shape, dtype=dcol.dtype)
    d = dcol.view(n, oh, ow, b.kp)
    for iy in range(h):
        vy = [iy * b.up]
        if b.reflect:
            if 1 <= iy <= b.pad:
                vy.append(-iy)
            if h - 1 - b.pad <= iy <= h - 2:
                vy.append(2 * (h - 1) - iy)
        for ix in range(w):
            vx = [ix * b.up]
            if b.reflect:
                if 1 <= ix <= b.pad:
                    vx.append(-ix)
                if w - 1 - b.pad <= ix <= w - 2:
                    vx.append(2 * (w - 1) - ix)
            for a in vy:
                for ky in range(b.k):
                    ty = a + b.pad - ky
                    if ty < 0 or ty % b.stride or ty // b.stride >= oh:
                        continue
                    for e in vx:
                        for kx in range(b.k):
                            tx = e + b.pad - kx
                            if tx < 0 or tx % b.stride or tx // b.stride >= ow:
                                continue
                            dx[:, iy, ix, :] += d[:, ty // b.stride, tx // b.stride,
                                                  (ky * b.k + kx) * c:(ky * b.k + kx + 1) * c]
    return dx


CASES = [
    ("conv3", lambda: nn.Conv2d(16, 32, 3, 1, 1), 0, 7),
    ("conv3 s2", lambda: nn.Conv2d(8, 16, 3, 2, 1), 0, 8),
    ("conv4 s2", lambda: nn.Conv2d(3, 16, 4, 2, 1), 0, 8),
    ("conv4 s1", lambda: nn.Conv2d(8, 3, 4, 1, 1), 0, 7),
    ("conv9 3ch", lambda: nn.Conv2d(3, 16, 9, 1, 4), 0, 10),
    ("conv1", lambda: nn.Conv2d(16, 1, 1, 1, 0), 0, 3),
    ("conv6 valid", lambda: nn.Conv2d(8, 16, 6, 1, 0), 0, 6),
    ("reflect3 conv7", lambda: nn.Conv2d(3, 16, 7, 1, 0), 3, 9),
    ("reflect1 conv3", lambda: nn.Conv2d(8, 8, 3, 1, 0), 1, 5),
    ("convT3 s2 op1", lambda: nn.ConvTranspose2d(16, 8, 3, 2, 1, output_padding=1), 0, 5),
]


@pytest.mark.parametrize("name,make,reflect,size", CASES, ids=[c[0] for c in CASES])
def test_conv_as_patch_gemm(name, make, reflect, size):
    from ipr_gan_b200 import seqnet
    torch.manual_seed(0)
    conv = make().double()
    b = seqnet.Block(conv, reflect)
    n, h, w = 2, size, size + 1
    x = torch.randn(n, conv.in_channels, h, w, dtype=torch.float64, requires_grad=True)
    ref = conv(F.pad(x, (reflect,) * 4, mode="reflect") if reflect else x)
    oh, ow = b.out_hw(h, w)
    assert tuple(ref.shape[2:]) == (oh, ow)
    # forward: patch matrix x packed weight (+ bias)
    xin = torch.zeros(n, h, w, b.cin_p, dtype=torch.float64)
    xin[..., :b.cin] = x.detach().permute(0, 2, 3, 1)
    col = emulate_im2col(xin, b, oh, ow)
    wmat = b.fwd_layout(conv.weight.detach())[0]                      # [n_p][kp]
    bl = b.bias_layout(torch.arange(b.cout, dtype=torch.float64))       # as PackSet sees it: source indices, -1 = zero
    bias = torch.where(bl >= 0, conv.bias.detach()[bl.clamp_min(0).long()], torch.zeros((), dtype=torch.float64))
    y = col @ wmat.t() + bias
    got = y.view(n, oh, ow, b.n_p)[..., :b.cout].permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=1e-10), name
    assert float(y.view(n, oh, ow, b.n_p)[..., b.cout:].abs().max()) == 0.0 if b.n_p > b.cout else True
    # backward
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dy0 = torch.zeros(n * oh * ow, b.n_p, dtype=torch.float64)
    dy0[:, :b.cout] = dy.permute(0, 2, 3, 1).reshape(-1, b.cout)
    dcol = dy0 @ b.dgrad_layout(conv.weight.detach())[0].t()          # [M][kp]
    dx = emulate_col2im(dcol, b, (n, h, w, b.cin_p), oh, ow)
    assert torch.allclose(dx[..., :b.cin].permute(0, 3, 1, 2), x.grad, atol=1e-10), name
    assert float(dx[..., b.cin:].abs().max()) == 0.0 if b.cin_p > b.cin else True
    # weight gradient: GEMM-layout dW[n][k] scattered through the tables
    off, s_n, row_map = b.wgrad_tables()
    g = dy0.t() @ col                                                 # [n_p][kp]
    dw = torch.zeros(conv.weight.numel(), dtype=torch.float64)
    hit = torch.zeros(conv.weight.numel(), dtype=torch.int64)
    for r in range(b.n_p):
        if int(row_map[r]) < 0:
            continue
        for k in range(b.kp):
            o = int(off[0, k])
            if o >= 0:
                dw[int(row_map[r]) * s_n + o] += g[r, k]
                hit[int(row_map[r]) * s_n + o] += 1
    assert int(hit.min()) == 1 and int(hit.max()) == 1               # every weight element written exactly once
    assert torch.allclose(dw.view_as(conv.weight), conv.weight.grad, atol=1e-9), name


def test_lowering_of_the_four_networks():
    import networks
    from ipr_gan_b200 import seqnet
    sr = seqnet.lower(networks.SRResNet())
    assert len(sr) == 37 and sr[0].act == seqnet.ACT_PRELU and sr[0].k == 9 and sr[-1].final and sr[-1].cout == 3
    assert [b.residual for b in sr[1:5]] == [None, 1, None, 3] and sr[33].residual == 1
    assert sr[34].shuffle and sr[35].shuffle and sr[34].cout == 256 and sr[34].act == seqnet.ACT_PRELU
    assert sum(b.norm is not None for b in sr) == 33
    d96 = seqnet.lower(networks.Discriminator96())
    assert [b.stride for b in d96] == [1, 2, 1, 2, 1, 2, 1, 2, 1, 1] and d96[8].k == 6 and d96[9].k == 1
    assert all(b.act == seqnet.ACT_LRELU and abs(b.slope - 0.2) < 1e-9 for b in d96[:9])
    rg = seqnet.lower(networks.Resnet9Blocks())
    assert len(rg) == 24 and rg[0].reflect and rg[0].pad == 3 and rg[-1].reflect and rg[-1].act == seqnet.ACT_TANH
    assert rg[21].transposed and rg[21].up == 2 and rg[21].pad == 1 and rg[21].out_hw(32, 32) == (64, 64)
    assert [rg[i].residual for i in (4, 6, 20)] == [3, 5, 19]
    assert sum(isinstance(b.norm, torch.nn.InstanceNorm2d) for b in rg) == 23
    pd = seqnet.lower(networks.ConvDiscriminator())
    assert [b.out_hw(128, 128) for b in pd[:1]] == [(64, 64)] and pd[3].stride == 1 and pd[4].cout == 1
    assert pd[1].norm.weight is None                                   # non-affine InstanceNorm
