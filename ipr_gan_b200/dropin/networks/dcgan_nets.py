"""DCGAN generator / discriminator with the reference's module tree, hence its ``state_dict`` keys:

  ConvGenerator     networks/conv_generator.py:3-33   fc.0.{weight,bias}, convs.{0,1,2}.0.weight,
                    convs.{0,1,2}.1.{weight,bias,running_mean,running_var,num_batches_tracked}, convs.3.weight
  SNDiscriminator   networks/sn_discriminator.py:4-38 net.{0,1,2}.{0,2}.{bias,weight_orig,weight_u,weight_v},
                    net.3.*, net.6.*   (legacy torch.nn.utils.spectral_norm state)

The child modules are ONLY the parameter / state containers: ``forward`` never calls them.  The whole network runs
as one autograd node over the library's implicit-GEMM (tcgen05) and normalisation kernels
(``ipr_gan_b200.engine``), NHWC/bf16 inside, NCHW/fp32 at the module boundary.  There is no CPU or PyTorch-op
path: a non-CUDA input raises (the modules can still be built, moved and (de)serialised on the CPU).
"""
import torch.nn as nn
from torch.nn.utils import spectral_norm as _sn


def _require_cuda(t, who):
    if not t.is_cuda:
        from ipr_gan_b200.ops import IprError
        raise IprError("%s.forward computes on CUDA only (libipr_b200.so, sm_100a); got a %s tensor -- move the "
                       "module and its input to a CUDA device" % (who, t.device.type))


class ConvGenerator(nn.Module):
    def __init__(self, mg, z_dim=128):
        super().__init__()
        self.mg = mg
        self.z_dim = z_dim
        widths = (512, 256, 128, 64)
        self.fc = nn.Sequential(nn.Linear(z_dim, widths[0] * mg * mg), nn.ReLU(inplace=True))
        stages = []
        for cin, cout in zip(widths[:-1], widths[1:]):
            stages.append(nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False),
                                        nn.BatchNorm2d(cout), nn.ReLU(inplace=True)))
        stages += [nn.ConvTranspose2d(widths[-1], 3, 3, 1, 1, bias=False), nn.Tanh()]
        self.convs = nn.Sequential(*stages)

    def forward(self, z):
        _require_cuda(z, "ConvGenerator")
        from ipr_gan_b200 import engine
        return engine.generator_forward(self, z)


class Flatten(nn.Module):
    def forward(self, x):
        return x.flatten(1)


class SNDiscriminator(nn.Module):
    def __init__(self, md):
        super().__init__()
        self.md = md

        def pair(cin, cout):
            return nn.Sequential(_sn(nn.Conv2d(cin, cout, 3, 1, 1, bias=True)), nn.LeakyReLU(0.1, inplace=True),
                                 _sn(nn.Conv2d(cout, cout, 4, 2, 1, bias=True)), nn.LeakyReLU(0.1, inplace=True))

        self.net = nn.Sequential(pair(3, 64), pair(64, 128), pair(128, 256),
                                 _sn(nn.Conv2d(256, 512, 3, 1, 1, bias=True)), nn.LeakyReLU(0.1, inplace=True),
                                 Flatten(), _sn(nn.Linear(512 * md * md, 1)))

    def forward(self, x):
        _require_cuda(x, "SNDiscriminator")
        from ipr_gan_b200 import engine
        return engine.discriminator_forward(self, x)


def ConvGenerator32():
    return ConvGenerator(mg=4)


def ConvGenerator64():
    return ConvGenerator(mg=8)


def SNDiscriminator32():
    return SNDiscriminator(md=4)


def SNDiscriminator64():
    return SNDiscriminator(md=8)
