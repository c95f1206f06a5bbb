"""Drop-in ``networks`` package: zero-argument factories looked up by name
(``getattr(networks, config.G)()``, models/dcgan.py:10-17) returning modules whose ``state_dict``
keys and shapes equal the reference's (checkpoint format).

Names outside the accelerated path (networks/__init__.py:4-6,10: the VAE encoder / decoder, the FID Inception
network, the VGG feature extractor) resolve lazily to the reference's OWN modules, executed in place from the
reference checkout (``ipr_gan_b200.refpath``) -- they stay plain PyTorch and nothing of them is copied here."""
from ipr_gan_b200 import refpath as _refpath
from networks.dcgan_nets import (ConvGenerator, ConvGenerator32, ConvGenerator64, SNDiscriminator,  # noqa: F401
                                 SNDiscriminator32, SNDiscriminator64)
from networks.seq_nets import (ConvDiscriminator, Discriminator96, Resnet6Blocks, Resnet9Blocks,  # noqa: F401
                                 ResnetGenerator, SRResNet, VGG19Feature)

__getattr__ = _refpath.passthrough("networks", {
    "InceptionActivations": "inception", "InceptionV3": "inception", "fid_inception_v3": "inception",
    "Encoder32": "encoder", "Encoder64": "encoder", "Decoder32": "decoder", "Decoder64": "decoder",
})
