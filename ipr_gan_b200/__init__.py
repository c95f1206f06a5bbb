"""ipr_gan_b200 -- B200-native (sm_100a) implementation of the IPR-GAN protected training step.

Layout
  csrc/ + libipr_b200.so   hand-written CUDA kernels behind the C ABI of include/ipr_b200.h
  _lib.py / ops.py         ctypes binding and tensor-level entry points (no CPU fallback)
  engine.py                whole-network autograd nodes (generator / discriminator) over the kernels
  dropin/                  packages named like the reference's (tools, models, networks, configs,
                           pytorch_msssim) with the same call signatures; put on sys.path by
                           ``enable_dropin()`` so the reference's train.py / eval.py / sign_flip.py
                           import them instead of its own.
"""
import os
import sys

__version__ = "0.1.0"

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DROPIN_DIR = os.path.join(PKG_DIR, "dropin")
_NAMES = ("tools", "models", "networks", "configs", "pytorch_msssim")


def enable_dropin(reference_root=None):
    """Make ``import tools, models, networks, configs, pytorch_msssim`` resolve to this package's
    drop-in modules.  Idempotent.  Raises if modules of those names were already imported from
    somewhere else (e.g. the reference tree) -- mixing the two silently would void parity claims.

    ``reference_root``: the ipr-gan checkout whose out-of-scope modules (``networks.InceptionActivations``,
    ``models.VAE`` ...) the drop-in packages pass through to (default: found on ``sys.path`` / the working
    directory / ``$IPR_REFERENCE_ROOT``, see ``ipr_gan_b200.refpath``)."""
    if reference_root is not None:
        from . import refpath
        refpath.set_reference_root(reference_root)
    root = os.path.dirname(PKG_DIR)
    for p in (root, DROPIN_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, root)
    sys.path.insert(0, DROPIN_DIR)
    for name in _NAMES:
        mod = sys.modules.get(name)
        if mod is not None:
            origin = os.path.abspath(getattr(mod, "__file__", "") or "")
            if not origin.startswith(DROPIN_DIR):
                raise RuntimeError("module %r already imported from %s; enable_dropin() must run first"
                                   % (name, origin))
    return DROPIN_DIR
