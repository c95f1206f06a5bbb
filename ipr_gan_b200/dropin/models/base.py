"""``models.base`` under its reference name (models/base.py:4-79): the reference's out-of-scope model files
(models/vae.py:2) import ``Model`` from here when they run over the drop-in packages."""
from models.core import Model, Wrapper  # noqa: F401
