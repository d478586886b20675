"""Inert stand-in."""
from . import axes_grid1  # noqa: F401
