"""causarray_b200 -- B200-native (sm_100a) GCATE + doubly-robust LFC hot path of causarray.

Drop-in for the reference's public API on that path (``causarray/__init__.py:10-32``):
``gcate_lfc_batch``, ``LFC``, ``fit_gcate``, ``fit_gcate_batch``, ``estimate_r``,
``alter_min``, ``fit_glm`` / ``fit_glm_auto`` / ``fit_glm_fast``, ``estimate_disp``,
``estimate_propensity_scores``, ``prep_causarray_data``, ``comp_size_factor`` ...
The numerics run in ``libcausarray_b200.so`` (hand-written CUDA, C ABI in
``include/causarray_b200.h``); there is no CPU fallback.
"""
from .prep import (prep_causarray_data, comp_size_factor, reset_random_seeds, Early_Stopping, _filter_params,
                   subsample_ctrl_cells, subsample_pert_cells)
from .optim import nll, grad, update, update_with_mask, alter_min
from .glm import (fit_glm, fit_glm_auto, fit_glm_fast, estimate_disp, estimate_disp_auto, estimate_disp_fast,
                  _backend_override)
from .estimators import fit_gcate, fit_gcate_batch, estimate, estimate_r
from .inference import bh_correction, fdx_control, multiplier_bootstrap, step_down, augmentation
from .dr import (LFC, gcate_lfc_batch, LFC_batch, compute_causal_estimand, cross_fitting, AIPW_mean,
                 estimate_propensity_scores)

__version__ = "0.1.0"

__all__ = [
    "LFC", "gcate_lfc_batch", "LFC_batch", "fit_glm", "fit_glm_fast", "fit_glm_auto", "reset_random_seeds",
    "fit_gcate", "fit_gcate_batch", "estimate_r", "estimate", "alter_min", "estimate_propensity_scores",
    "prep_causarray_data", "comp_size_factor", "compute_causal_estimand", "cross_fitting", "AIPW_mean",
    "nll", "grad", "update", "update_with_mask", "Early_Stopping", "bh_correction", "fdx_control",
    "estimate_disp", "estimate_disp_auto", "estimate_disp_fast", "subsample_ctrl_cells", "subsample_pert_cells",
]
