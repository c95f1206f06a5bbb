"""Drop-in replacement of the reference's ``tools`` package (tools/__init__.py:1-8): same names,
same call signatures, computing on libipr_b200.so."""
from tools.losses import *                                            # noqa: F401,F403
from tools.signature import SignLossModel, BitGenerator               # noqa: F401
from tools.triggers import (PasteWatermark, RandomBitMask, RandomNoisePatch, TransformDist,  # noqa: F401
                            TransformVar)
from tools.verify import compute_matching_prob, compute_hash          # noqa: F401
