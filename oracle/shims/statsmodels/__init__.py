"""Stand-in for statsmodels (absent from the image): see oracle/shims/README.md."""
from . import api  # noqa: F401
from . import stats  # noqa: F401
