"""SRGAN / CycleGAN networks with the reference's module trees (identical ``state_dict`` keys and initialisation):

  SRResNet           networks/sr_resnet.py:3-44          Discriminator96     networks/discriminator_96.py:3-35
  Resnet{6,9}Blocks  networks/resnet_generator.py:3-59   ConvDiscriminator   networks/conv_discriminator.py:3-21
  VGG19Feature       networks/vgg.py:5-40 (frozen extractor; out of the accelerated scope)

As with the DCGAN networks the child modules are ONLY the parameter / state containers: ``forward`` hands the module
tree to ``ipr_gan_b200.seqnet``, which lowers it to conv / norm / activation blocks and runs the whole network as one
autograd node over the library's kernels -- every convolution (k 1..9, stride 1/2, reflection or zero border,
transposed) on the tcgen05 GEMM through a patch matrix, BatchNorm / InstanceNorm / PReLU / PixelShuffle in
csrc/layers.cu, NHWC bf16 inside, NCHW fp32 at the module boundary.  No CPU or PyTorch-op path: a non-CUDA input
raises (``torch.nn.Sequential.forward(net, x)`` still evaluates the same tree with PyTorch operators; the parity
tests use exactly that as the fp32 reference).
"""
import torch.nn as nn


def _native(net, x, who):
    if not x.is_cuda:
        from ipr_gan_b200.ops import IprError
        raise IprError("%s.forward computes on CUDA only (libipr_b200.so, sm_100a); got a %s tensor -- move the "
                       "module and its input to a CUDA device" % (who, x.device.type))
    from ipr_gan_b200 import seqnet
    return seqnet.forward(net, x)


# ------------------------------------------------------------------------------------------------ SRResNet
class _SRConv(nn.Sequential):
    def __init__(self, cin, cout, k, s=1, p=0, n=False, a=None):
        layers = [nn.Conv2d(cin, cout, k, s, p)]
        if n:
            layers.append(nn.BatchNorm2d(cout))
        if a:
            layers.append(a)
        super().__init__(*layers)
        nn.init.kaiming_normal_(self[0].weight.data, a=0.25 if a else 1.0, mode="fan_in")
        self[0].bias.data.zero_()


class _Skip(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.block = block

    def forward(self, x):
        return x + self.block(x)


class _Up2(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(_SRConv(cin, cout * 4, 3, 1, 1), nn.PixelShuffle(2), nn.PReLU())


class SRResNet(nn.Sequential):
    _ipr_native_norms = True            # BatchNorm backward runs in csrc/layers.cu (sign-loss gradient fused there)

    def __init__(self, n_block=16):
        trunk = [_Skip(nn.Sequential(_SRConv(64, 64, 3, 1, 1, n=True, a=nn.PReLU()), _SRConv(64, 64, 3, 1, 1, n=True)))
                 for _ in range(n_block)]
        trunk.append(_SRConv(64, 64, 3, 1, 1, n=True))
        super().__init__(_SRConv(3, 64, 9, 1, 4, a=nn.PReLU()), _Skip(nn.Sequential(*trunk)),
                         _Up2(64, 64), _Up2(64, 64), _SRConv(64, 3, 9, 1, 4))

    def forward(self, x):
        return _native(self, x, "SRResNet")


# ------------------------------------------------------------------------------------------------ Discriminator96
class _DConv(nn.Sequential):
    def __init__(self, cin, cout, k, s=1, p=0):
        super().__init__(nn.Conv2d(cin, cout, k, s, p), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, True))
        nn.init.kaiming_normal_(self[0].weight.data, a=0.2, mode="fan_in")
        self[0].bias.data.zero_()


class Discriminator96(nn.Sequential):
    def __init__(self):
        widths = [(64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2), (256, 512, 1), (512, 512, 2)]
        super().__init__(nn.Conv2d(3, 64, 3, 1, 1), nn.LeakyReLU(0.2, True),
                         *[_DConv(ci, co, 3, s, 1) for ci, co, s in widths],
                         nn.Conv2d(512, 1024, 6, 1, 0), nn.LeakyReLU(0.2, True), nn.Conv2d(1024, 1, 1, 1, 0))

    def forward(self, x):
        return _native(self, x, "Discriminator96").squeeze()


# ------------------------------------------------------------------------------------------------ CycleGAN
class ResnetBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(ch, ch, 3, 1, 0, bias=True),
                                   nn.InstanceNorm2d(ch, affine=True), nn.ReLU(True),
                                   nn.ReflectionPad2d(1), nn.Conv2d(ch, ch, 3, 1, 0, bias=True),
                                   nn.InstanceNorm2d(ch, affine=True))

    def forward(self, x):
        return x + self.block(x)


class ResnetGenerator(nn.Sequential):
    _ipr_native_norms = True            # InstanceNorm backward runs in csrc/layers.cu (sign-loss gradient fused there)

    def __init__(self, n_block):
        seq = [nn.ReflectionPad2d(3), nn.Conv2d(3, 64, 7, 1, 0), nn.InstanceNorm2d(64, affine=True), nn.ReLU(True)]
        for ch in (64, 128):
            seq += [nn.Conv2d(ch, ch * 2, 3, 2, 1), nn.InstanceNorm2d(ch * 2, affine=True), nn.ReLU(True)]
        seq += [ResnetBlock(256) for _ in range(n_block)]
        for ch in (256, 128):
            seq += [nn.ConvTranspose2d(ch, ch // 2, 3, 2, 1, output_padding=1), nn.InstanceNorm2d(ch // 2, affine=True),
                    nn.ReLU(True)]
        seq += [nn.ReflectionPad2d(3), nn.Conv2d(64, 3, 7, 1, 0), nn.Tanh()]
        super().__init__(*seq)

    def forward(self, x):
        return _native(self, x, "ResnetGenerator")


def Resnet9Blocks():
    return ResnetGenerator(n_block=9)


def Resnet6Blocks():
    return ResnetGenerator(n_block=6)


class ConvDiscriminator(nn.Sequential):
    def __init__(self):
        super().__init__(nn.Conv2d(3, 64, 4, 2, 1), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(64, 128, 4, 2, 1), nn.InstanceNorm2d(128), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(128, 256, 4, 2, 1), nn.InstanceNorm2d(256), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(256, 512, 4, 1, 1), nn.InstanceNorm2d(512), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(512, 1, 4, 1, 1))

    def forward(self, x):
        return _native(self, x, "ConvDiscriminator")


# ------------------------------------------------------------------------------------------------ VGG feature net
class VGG19Feature(nn.Module):
    """Frozen VGG-19 feature extractor up to ``layer`` (networks/vgg.py).  The reference loads ImageNet weights from
    the network; offline (``IPR_VGG_RANDOM_INIT=1`` or no weights available) the seeded random initialisation is used."""
    _order = ("conv1_1 relu1_1 conv1_2 relu1_2 pool1 conv2_1 relu2_1 conv2_2 relu2_2 pool2 conv3_1 relu3_1 conv3_2 relu3_2 "
              "conv3_3 relu3_3 conv3_4 relu3_4 pool3 conv4_1 relu4_1 conv4_2 relu4_2 conv4_3 relu4_3 conv4_4 relu4_4 pool4 "
              "conv5_1 relu5_1 conv5_2 relu5_2 conv5_3 relu5_3 conv5_4 relu5_4 pool5").split()

    def __init__(self, layer="relu5_4"):
        super().__init__()
        import os
        from torchvision.models import vgg19
        cut = self._order.index(layer) + 1
        try:
            if os.environ.get("IPR_VGG_RANDOM_INIT"):
                raise RuntimeError("random init requested")
            full = vgg19(weights="IMAGENET1K_V1")
        except Exception:
            full = vgg19(weights=None)
        self.net = full.features[:cut]
        self.net.eval()
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.net(x)
