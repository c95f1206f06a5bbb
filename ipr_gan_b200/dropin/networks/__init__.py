"""Drop-in ``networks`` package: zero-argument factories looked up by name
(``getattr(networks, config.G)()``, models/dcgan.py:10-17) returning modules whose ``state_dict``
keys and shapes equal the reference's (checkpoint format)."""
from networks.dcgan_nets import (ConvGenerator, ConvGenerator32, ConvGenerator64, SNDiscriminator,  # noqa: F401
                                 SNDiscriminator32, SNDiscriminator64)
from networks.torch_nets import (ConvDiscriminator, Discriminator96, Resnet6Blocks, Resnet9Blocks,  # noqa: F401
                                 ResnetGenerator, SRResNet, VGG19Feature)
