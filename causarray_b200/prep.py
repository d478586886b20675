"""Host-side data preparation mirrored from the reference's ``causarray/utils.py``.

Same names, arguments and results as the reference (``prep_causarray_data``
``:19-65``, ``comp_size_factor`` ``:179-219``, ``Early_Stopping`` ``:101-161``,
``_filter_params`` ``:74-98``, the control / perturbation subsamplers
``:223-278``).  These are O(n p) host passes executed once per data set; the
size-factor median is row SURVEY.md section 8(f).1 ("next" on the device).
"""
import inspect
import os
import random

import numpy as np
import pandas as pd


def reset_random_seeds(seed):
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    random.seed(seed)


def _filter_params(func, kwargs):
    """Keep the entries of ``kwargs`` that ``func`` (a callable or a dict of names) accepts."""
    if isinstance(func, dict):
        names = func.keys()
    elif callable(func):
        names = inspect.signature(func).parameters.keys()
    else:
        raise ValueError("The provided func is not a callable function, or a dict.")
    return {k: v for k, v in kwargs.items() if k in names}


def prep_causarray_data(Y, A, X=None, X_A=None, intercept=True):
    """Cap extreme counts, add the intercept and the standardised log library size.

    Returns ``(Y, A, X, X_A)`` exactly as the reference does: counts are capped at
    the rounded 99.9 % quantile of the per-gene maxima, ``X`` gains a leading
    column of ones, ``X_A`` gains the intercept and a trailing z-scored
    ``log2`` library size.
    """
    if not isinstance(Y, pd.DataFrame):
        Y = np.asarray(Y)
    cap = np.round(np.quantile(np.max(Y, 0), 0.999))
    Y = np.minimum(Y, cap)
    if not isinstance(A, pd.DataFrame):
        A = np.asarray(A)
    if A.ndim == 1:
        A = A[:, None]
    n = Y.shape[0]
    X = np.zeros((n, 0)) if X is None else np.asarray(X)
    X_A = X if X_A is None else np.asarray(X_A)
    lib = np.log2(np.sum(np.asarray(Y), axis=1))
    lib = (lib - np.mean(lib)) / np.std(lib, ddof=1)
    X_A = np.hstack((X_A, lib[:, None]))
    ones = np.ones((n, 1)) if intercept else np.empty((n, 0))
    return Y, A, np.hstack((ones, X)), np.hstack((ones, X_A))


def comp_size_factor(counts, method='geomeans', lib_size=1e4, **kwargs):
    """Median-of-ratios size factors (``method='geomeans'``) or library scaling (``'scale'``)."""
    if method == 'geomeans':
        counts = np.asarray(counts, dtype=np.float64)
        pos = counts > 0
        with np.errstate(divide='ignore', invalid='ignore'):
            logc = np.where(pos, np.log(counts), np.nan)
            log_geo = np.nanmean(logc, axis=0)
        log_geo[~pos.any(axis=0)] = -np.inf
        ok = np.isfinite(log_geo)[None, :] & pos
        ratio = np.where(ok, logc - log_geo[None, :], np.nan)
        log_sf = np.nanmedian(ratio, axis=1)
        return np.exp(log_sf - np.mean(log_sf))
    if method == 'scale':
        return 1. / np.sum(counts, axis=0) * lib_size
    raise ValueError("Method must be in {'geomeans' or 'scale'}.")


class Early_Stopping():
    """Early-stopping monitor: progress means improving the best metric by more than
    ``tolerance + rel_tol * |best|``; stop once ``patience`` steps pass without progress."""

    def __init__(self, warmup=25, patience=25, tolerance=0., rel_tol=1e-4, is_minimize=True, **kwargs):
        self.warmup, self.patience = warmup, patience
        self.tolerance, self.rel_tol = tolerance, rel_tol
        self.is_minimize = is_minimize
        self.factor = 1.0 if is_minimize else -1.0
        self.step = -1
        self.best_step = -1
        self.best_metric = np.inf
        self.info = None

    def __call__(self, metric):
        self.step += 1
        if self.step < self.warmup:
            return False
        thr = self.tolerance + self.rel_tol * abs(self.best_metric) if np.isfinite(self.best_metric) else 0.
        if self.factor * metric < self.factor * self.best_metric - thr:
            self.best_metric, self.best_step = metric, self.step
            return False
        if self.step - self.best_step > self.patience:
            self.info = 'Best Epoch: %d. Best Metric: %f.' % (self.best_step, self.best_metric)
            return True
        return False

    def reset_state(self):
        self.best_step = self.step
        self.best_metric = np.inf


def subsample_ctrl_cells(ctrl_idx, n_ctrl=2000, random_state=0):
    """Fixed control subsample shared by every perturbation batch (sorted indices)."""
    ctrl_idx = np.asarray(ctrl_idx)
    if len(ctrl_idx) <= n_ctrl:
        return ctrl_idx
    rng = np.random.default_rng(random_state)
    return np.sort(rng.choice(ctrl_idx, size=n_ctrl, replace=False))


def subsample_pert_cells(pert_idx, max_cells=2000, random_state=0):
    """Cap the perturbed cells of one batch at ``max_cells`` (sorted indices)."""
    pert_idx = np.asarray(pert_idx)
    if max_cells is None or len(pert_idx) <= max_cells:
        return pert_idx
    rng = np.random.default_rng(random_state)
    return np.sort(rng.choice(pert_idx, size=max_cells, replace=False))
