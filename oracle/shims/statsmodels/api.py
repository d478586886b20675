"""``statsmodels.api`` names used by the reference (``causarray/gcate_glm.py:187-236``)."""
from oracle.glm import SMGLM as GLM  # noqa: F401
from oracle.glm import families  # noqa: F401

OLS = None
