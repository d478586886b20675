"""Inert stand-in."""
