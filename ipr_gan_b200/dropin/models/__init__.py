"""Drop-in ``models`` package (models/__init__.py:1-6)."""
from models.core import Model, Wrapper                                  # noqa: F401
from models.cyclegan import CycleGAN, ImagePool                         # noqa: F401
from models.dcgan import DCGAN                                          # noqa: F401
from models.protect import BlackBoxWrapper, WhiteBoxWrapper             # noqa: F401
from models.srgan import SRGAN                                          # noqa: F401
from models.util import DisableBatchNormStats, Replica                  # noqa: F401
