"""Pass-through to the reference tree for everything OUTSIDE the accelerated path.

The drop-in packages (``networks``, ``models``, ``tools``) replace only what SURVEY.md section 8 puts on the hot
path.  The reference's entry scripts still import names that are out of scope here -- ``networks.InceptionActivations``
(FID / IS features, experiments/image_generation.py:5), ``networks.Encoder*/Decoder*`` and ``models.VAE`` (the VAE
baseline, models/__init__.py:5), ``networks.VGG19Feature`` (SRGAN content loss).  Those keep running as the
reference's OWN PyTorch code: a drop-in package resolves such a name lazily (module ``__getattr__``, PEP 562) by
executing the reference's source file in place, under the drop-in package's namespace, so its ``import networks`` /
``from models.base import Model`` land on the drop-in modules.  Nothing is copied; without a reference tree the
names raise ``AttributeError`` with an explanation.

The reference root is, in order: ``$IPR_REFERENCE_ROOT``, a root passed to ``enable_dropin(reference_root=...)``,
or the first ``sys.path`` entry / the working directory that looks like the reference checkout.
"""
import importlib.util
import os
import sys

_ROOT = None
_MARKERS = (os.path.join("experiments", "image_generation.py"), os.path.join("networks", "inception.py"))


def _looks_like_reference(path):
    return bool(path) and all(os.path.isfile(os.path.join(path, m)) for m in _MARKERS)


def set_reference_root(path):
    global _ROOT
    if path is not None and not _looks_like_reference(path):
        raise RuntimeError("%r is not an ipr-gan checkout (experiments/image_generation.py, networks/inception.py)" % path)
    _ROOT = os.path.abspath(path) if path else None


def reference_root():
    if _ROOT:
        return _ROOT
    env = os.environ.get("IPR_REFERENCE_ROOT")
    if _looks_like_reference(env):
        return os.path.abspath(env)
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in [os.getcwd()] + list(sys.path):
        cand = os.path.abspath(cand or os.getcwd())
        if cand.startswith(here):
            continue
        if _looks_like_reference(cand):
            return cand
    return None


def load(package, submodule):
    """Execute ``<reference>/<package>/<submodule>.py`` as module ``<package>.<submodule>`` (once) and return it."""
    name = "%s.%s" % (package, submodule)
    mod = sys.modules.get(name)
    if mod is not None:
        return mod
    root = reference_root()
    if root is None:
        raise AttributeError(
            "%s is outside the accelerated path and is served by the reference's own %s/%s.py, but no reference "
            "checkout was found (set IPR_REFERENCE_ROOT or run from the reference directory)" % (name, package, submodule))
    path = os.path.join(root, package, submodule + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    return mod


def passthrough(package, table):
    """-> a module-level ``__getattr__`` resolving the names in ``table`` ({name: reference submodule})."""
    def __getattr__(name):
        sub = table.get(name)
        if sub is None:
            raise AttributeError("module %r has no attribute %r" % (package, name))
        value = getattr(load(package, sub), name)
        sys.modules[package].__dict__[name] = value
        return value
    return __getattr__
