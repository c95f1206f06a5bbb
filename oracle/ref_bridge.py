"""oracle/ref_bridge.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference (``/root/reference``) on CPU so that the restated oracle
(``oracle/ipr_oracle.py``) can be validated against it and golden vectors can be generated
(``oracle/make_golden.py``).  ``/root/reference`` exists only in the build container, never on
the GPU box, so nothing that runs under ``-m gpu``, ``smoke()`` or ``bench.py`` imports this file.

Shims needed for the reference to import under torch 2.11 / NumPy 2:
* ``pytorch_msssim``  -> oracle/shims/pytorch_msssim  (restatement, package absent)
* ``pdqhash``         -> oracle/shims/pdqhash         (C restatement, package absent)
* ``np.bool8``        -> alias of ``np.bool_``        (removed in NumPy 2; tools/phash_pvalue.py:14)
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("IPR_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tools"))


_CLASH = ("tools", "models", "networks", "configs", "experiments", "datasets", "pytorch_msssim", "pdqhash")


class reference_modules:
    """Context manager: inside it ``import tools, models, networks, configs`` resolve to the
    reference's packages; on exit the previous ``sys.modules`` / ``sys.path`` entries are restored
    so the product's drop-in packages of the same names are not shadowed."""

    def __enter__(self):
        import numpy as np
        if not hasattr(np, "bool8"):
            np.bool8 = np.bool_
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        self._saved_path = list(sys.path)
        self._saved_mods = {k: v for k, v in sys.modules.items()
                            if k.split(".")[0] in _CLASH}
        for k in list(self._saved_mods):
            del sys.modules[k]
        sys.path.insert(0, REFERENCE_ROOT)
        sys.path.insert(0, _SHIMS)
        mods = {}
        for name in ("configs", "networks", "tools", "models"):
            mods[name] = importlib.import_module(name)
        self.mods = mods
        return mods

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split(".")[0] in _CLASH]:
            del sys.modules[k]
        sys.modules.update(self._saved_mods)
        sys.path[:] = self._saved_path
        return False
