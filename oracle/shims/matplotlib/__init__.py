"""Inert stand-in (plots are out of scope)."""
from . import pyplot  # noqa: F401
