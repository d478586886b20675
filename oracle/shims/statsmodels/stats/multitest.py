"""``statsmodels.stats.multitest`` names used by the reference (``causarray/DR_inference.py:186``)."""
from oracle.inference import multipletests, fdrcorrection  # noqa: F401
