"""``models.wrappers`` under its reference name (models/wrappers.py:7-125)."""
from models.protect import BlackBoxWrapper, WhiteBoxWrapper  # noqa: F401
