"""Drop-in ``models`` package (models/__init__.py:1-6).  ``VAE`` (a baseline outside the accelerated path) resolves
lazily to the reference's own models/vae.py executed over these packages (``ipr_gan_b200.refpath``)."""
from ipr_gan_b200 import refpath as _refpath
from models.core import Model, Wrapper                                  # noqa: F401
from models.cyclegan import CycleGAN, ImagePool                         # noqa: F401
from models.dcgan import DCGAN                                          # noqa: F401
from models.protect import BlackBoxWrapper, WhiteBoxWrapper             # noqa: F401
from models.srgan import SRGAN                                          # noqa: F401
from models.util import DisableBatchNormStats, Replica                  # noqa: F401

__getattr__ = _refpath.passthrough("models", {"VAE": "vae"})
