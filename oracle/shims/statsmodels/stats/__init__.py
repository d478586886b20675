from . import multitest  # noqa: F401
