"""Doubly-robust (AIPW) log-fold-change estimation -- the S3 seam of SURVEY.md section 8(b).

Interface of the reference's ``causarray/DR_estimation.py`` (``AIPW_mean``
``:624-659``, ``cross_fitting`` ``:472-618``, ``estimate_propensity_scores``
``:66-200``) and ``causarray/DR_learner.py`` (``compute_causal_estimand``
``:32-269``, ``LFC`` ``:272-493``, ``gcate_lfc_batch`` ``:584-866``, ``LFC_batch``
``:869-883``): same arguments, defaults, DataFrame columns and error behaviour.

What differs by design: the (n, p, a, 2) tensors ``Y_hat`` and the pseudo-outcomes
are never materialised on the fast path.  The outcome model is one joint GLM per
gene, so ``Yhat_0 = exp(W b_cov + offset)`` and ``Yhat_1[k] = Yhat_0 * exp(b_trt[k])``;
``ca_aipw_lfc`` recomputes them inside the fused reduction (``csrc/dr.cu``).
``estimation['Y_hat']`` is therefore a lazy ``YhatFactors`` object (call
``.materialize()`` for the dense tensor).  A user-supplied dense ``Y_hat`` goes
through ``ca_aipw_lfc_yhat``.  Propensity scores stay on the host (sklearn
lbfgs, as in the reference -- SURVEY.md section 2, row 7).
"""
import contextlib
import gc
import os
import sys
import time
import warnings
from typing import Literal

import numpy as np
import pandas as pd
from sklearn.linear_model import LogisticRegression
from sklearn.model_selection import KFold

from . import glm as _glm_mod
from .device import RowsView, GlmDevice, aipw_lfc, aipw_lfc_yhat
from .inference import bh_correction, bh_correction_device, bh_correction_multi, fdx_control
from .prep import comp_size_factor, _filter_params, reset_random_seeds
from ._timing import phase

__version__ = "0.1.0"

import concurrent.futures as _cf
_PS_POOL = _cf.ThreadPoolExecutor(max_workers=1, thread_name_prefix='causarray_ps')


class YhatFactors:
    """Factorised counterfactual predictions of the joint outcome GLM."""

    def __init__(self, W, Bcov, Btrt, offset, cap=1e5):
        self.W, self.Bcov, self.Btrt, self.offset, self.cap = W, Bcov, Btrt, offset, cap
        self.shape = (W.shape[0], Bcov.shape[0], Btrt.shape[1], 2)

    def materialize(self, dtype=np.float64):
        eta0 = self.W @ self.Bcov.T
        if self.offset is not None:
            eta0 = eta0 + self.offset[:, None]
        out = np.empty(self.shape, dtype=dtype)
        with np.errstate(over='ignore'):
            out[:, :, :, 0] = np.minimum(np.exp(eta0), self.cap)[:, :, None]
            out[:, :, :, 1] = np.minimum(np.exp(eta0[:, :, None] + self.Btrt[None, :, :]), self.cap)
        return out


def AIPW_mean(Y, A, mu, pi):
    """AIPW pseudo-outcomes ``A/pi (Y - mu) + mu`` and their float64 means.

    ``Y (n, p)``, ``A, pi (n, a, 2)``, ``mu (n, p, a, 2)`` -> ``tau (p, a, 2)``,
    ``pseudo_y (n, p, a, 2)``.  Compatibility helper for small inputs (it materialises
    the 4-D tensor); ``LFC`` uses the fused device reduction instead.
    """
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        w = (A / pi)[:, None, ...]
    pseudo_y = w * (np.asarray(Y)[:, :, None, None] - mu) + mu
    return np.mean(pseudo_y, axis=0, dtype=np.float64), pseudo_y


def _validate_clip(clip):
    msg = 'clip must be None or a pair 0 <= lower < upper <= 1'
    if isinstance(clip, (str, bytes)) or np.ndim(clip) != 1:
        raise ValueError(msg)
    try:
        lo, hi = (float(b) for b in clip)
    except (TypeError, ValueError):
        raise ValueError(msg) from None
    if not 0 <= lo < hi <= 1:
        raise ValueError(msg)
    return lo, hi




_LEAN_CLOSS = None
_PS_LOSS_THREADS = max(1, int(os.environ.get('CAUSARRAY_B200_PS_THREADS', '2')))


def _fit_logistic_lean(x, y, x_te, C, balanced, tol, max_iter):
    """sklearn's binary ``LogisticRegression(solver='lbfgs', fit_intercept=False)`` without the
    estimator plumbing: the SAME pointwise loss kernel (``HalfBinomialLoss().closs``), sample
    weights, start vector, objective arithmetic (``LinearModelLoss.loss_gradient``:
    ``sum(loss) / sw_sum + l2 / 2 w.w``, gradient ``X^T (g / sw_sum) + l2 w``) and
    ``scipy.optimize.minimize(method='L-BFGS-B')`` options as
    ``sklearn/linear_model/_logistic.py::_logistic_regression_path`` -- identical function values, so
    identical iterates and (bit for bit) identical probabilities
    (``tests/test_host_api.py::test_lean_logistic_is_sklearn``) -- at a fraction of the cost per fit
    (input validation, label encoding, parameter routing and array-API dispatch dominate sklearn's
    5 ms on these 2 000 x 12 designs).  The reference's answer is defined by this solver and its
    ``tol=1e-4`` stop (``causarray/DR_estimation.py:18-42``), which is why the fit stays on the host."""
    global _LEAN_CLOSS
    import scipy.optimize
    from scipy.special import expit
    if _LEAN_CLOSS is None:
        from sklearn._loss.loss import HalfBinomialLoss
        _LEAN_CLOSS = HalfBinomialLoss().closs
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    pos = np.asarray(y) == 1
    target = np.ones(n)
    target[~pos] = 0.0
    if balanced:
        cnt = np.array([n - pos.sum(), pos.sum()], dtype=np.float64)
        cw = cnt.sum() / (2 * cnt)
        sw = np.ones(n) * cw[pos.astype(np.intp)]
        sw_sum = float(np.sum(sw))
    else:
        sw, sw_sum = None, n
    l2 = 1.0 / (C * sw_sum)
    xt = x.T
    loss_out, grad_out = np.empty(n), np.empty(n)
    closs = _LEAN_CLOSS
    # the pointwise kernel is 70 % of an objective evaluation (scalar exp / log1p per cell); its elements are
    # independent, so two OpenMP threads return the same arrays bit for bit in half the time
    n_thr = _PS_LOSS_THREADS

    def func(w):
        raw = x @ w + 0.0
        closs.loss_gradient(y_true=target, raw_prediction=raw, sample_weight=sw, loss_out=loss_out,
                            gradient_out=grad_out, n_threads=n_thr)
        loss = float(np.sum(loss_out) / sw_sum)
        loss += 0.5 * l2 * (w @ w)
        gp = grad_out / sw_sum
        return loss, xt @ gp + l2 * w

    w_hat = _lbfgsb_minimize(func, np.zeros(x.shape[1]), max_iter, 50, tol, 64 * np.finfo(float).eps)
    return expit((np.asarray(x_te, dtype=np.float64) @ w_hat.reshape(1, -1).T).ravel())


_LBFGSB_DIRECT = None    # None: not probed yet; True / False: the direct driver reproduces scipy.optimize.minimize


def _lbfgsb_direct(func, x0, maxiter, maxls, gtol, ftol, maxcor=10, maxfun=15000):
    """``scipy.optimize.minimize(method='L-BFGS-B', jac=True)`` without its Python plumbing: the same
    compiled ``setulb`` reverse-communication routine driven by the same loop
    (``scipy/optimize/_lbfgsb_py.py::_minimize_lbfgsb``), minus ``ScalarFunction`` / ``OptimizeResult`` /
    the callback hooks -- about half of the 1.3 ms a 2 000 x 12 fit took, and with 15 fits per batch the
    difference between hiding the propensity model behind the outcome GLM and waiting for it."""
    from scipy.optimize import _lbfgsb
    from scipy.optimize._lbfgsb_py import HAS_ILP64
    n = x0.shape[0]
    m = maxcor
    int_dtype = np.int64 if HAS_ILP64 else np.int32
    nbd = np.zeros(n, dtype=int_dtype)
    low_bnd = np.zeros(n, dtype=np.float64)
    upper_bnd = np.zeros(n, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros((n,), dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=int_dtype)
    task = np.zeros(2, dtype=int_dtype)
    ln_task = np.zeros(2, dtype=int_dtype)
    lsave = np.zeros(4, dtype=int_dtype)
    isave = np.zeros(44, dtype=int_dtype)
    dsave = np.zeros(29, dtype=np.float64)
    factr = ftol / np.finfo(float).eps
    n_iterations = nfev = 0
    while True:
        g = np.asarray(g, dtype=np.float64)
        _lbfgsb.setulb(m, x, low_bnd, upper_bnd, nbd, f, g, factr, gtol, wa, iwa, task, lsave, isave, dsave, maxls,
                       ln_task)
        if task[0] == 3:
            f, g = func(x.copy())
            nfev += 1
        elif task[0] == 1:
            n_iterations += 1
            if n_iterations >= maxiter:
                task[0] = 5
                task[1] = 504
            elif nfev > maxfun:
                task[0] = 5
                task[1] = 502
        else:
            break
    return x


def _lbfgsb_minimize(func, x0, maxiter, maxls, gtol, ftol):
    """The direct driver when it is PROVEN equal to scipy's on this installation (the first fit of a process is
    run through both and compared bit for bit -- ``setulb`` is a private interface), ``minimize`` otherwise."""
    global _LBFGSB_DIRECT
    import scipy.optimize

    def via_scipy():
        return scipy.optimize.minimize(func, x0, method='L-BFGS-B', jac=True,
                                       options={'maxiter': maxiter, 'maxls': maxls, 'gtol': gtol, 'ftol': ftol}).x
    if _LBFGSB_DIRECT is None or os.environ.get('CAUSARRAY_B200_PS_CHECK'):
        ref = via_scipy()
        try:
            mine = _lbfgsb_direct(func, x0, maxiter, maxls, gtol, ftol)
            same = bool(mine.shape == ref.shape and np.array_equal(mine, ref))
        except Exception:  # noqa: BLE001  (another scipy: different private signature)
            same = False
        if os.environ.get('CAUSARRAY_B200_PS_CHECK') and not same:
            raise AssertionError('direct L-BFGS-B driver differs from scipy.optimize.minimize')
        if _LBFGSB_DIRECT is None:
            _LBFGSB_DIRECT = same
        return ref
    if _LBFGSB_DIRECT:
        return _lbfgsb_direct(func, x0, maxiter, maxls, gtol, ftol)
    return via_scipy()


_LEAN_KEYS = {'fit_intercept', 'C', 'class_weight', 'random_state', 'tol', 'max_iter'}


def _fit_logistic(lr_kw, x, y, x_te):
    """One propensity fit: the lean path when the options are the plain L2 / lbfgs ones the reference
    uses, sklearn's estimator for anything else."""
    if (set(lr_kw) <= _LEAN_KEYS and lr_kw.get('fit_intercept') is False
            and (lr_kw.get('class_weight') is None or lr_kw.get('class_weight') == 'balanced')
            and not os.environ.get('CAUSARRAY_B200_PS_SKLEARN')):
        try:
            return _fit_logistic_lean(x, y, x_te, float(lr_kw.get('C', 1.0)), lr_kw.get('class_weight') == 'balanced',
                                      float(lr_kw.get('tol', 1e-4)), int(lr_kw.get('max_iter', 100)))
        except ImportError:      # sklearn moved its private loss modules: use the estimator
            pass
    return LogisticRegression(**lr_kw).fit(x, y).predict_proba(x_te)[:, 1]


def estimate_propensity_scores(A, X_A, K=1, ps_model='logistic', mask=None, clip=None, random_state=0,
                               verbose=False, class_weight='balanced', **kwargs):
    """Per-treatment logistic propensity scores against the shared all-zero control group.

    Logistic model: no intercept, ``C = 1``, ``class_weight`` (default ``'balanced'``), fitted on
    control plus case cells of each treatment; constant designs short-circuit to the class
    prior.  ``K > 1`` yields out-of-fold predictions.
    """
    A = np.asarray(A)
    if A.ndim == 1:
        A = A[:, None]
    X_A = np.asarray(X_A)
    if A.ndim != 2 or X_A.ndim != 2 or A.shape[0] != X_A.shape[0]:
        raise ValueError('A and X_A must be two-dimensional with matching rows')
    if not np.all(np.isin(A, (0, 1))):
        raise ValueError('A must contain only binary treatment indicators')
    try:
        K_int = int(K)
    except (TypeError, ValueError) as exc:
        raise ValueError('K must be a positive integer') from exc
    if K_int != K or K_int < 1:
        raise ValueError('K must be a positive integer')
    K = K_int
    if K > A.shape[0]:
        raise ValueError('K cannot exceed the number of samples')
    if mask is not None:
        mask = np.asarray(mask, dtype=bool)
        if mask.ndim == 1:
            mask = mask[:, None]
        if mask.shape != A.shape:
            raise ValueError('Mask must have the same shape as the treatment matrix')
    if ps_model != 'logistic':
        raise ValueError('Invalid propensity score model.' if ps_model not in ('random_forest_cv', 'ensemble')
                         else 'only the logistic propensity model is part of the accelerated path')
    lr_kw = {**{'fit_intercept': False, 'C': 1e0, 'class_weight': class_weight, 'random_state': 0}, **kwargs}
    lr_kw = _filter_params(LogisticRegression().get_params(), lr_kw)
    n = A.shape[0]
    if K == 1:
        folds = [(np.arange(n), np.arange(n))]
    elif K == n:
        folds = [(np.delete(np.arange(n), j), np.array([j])) for j in range(n)]
    else:
        folds = KFold(n_splits=K, random_state=random_state, shuffle=True).split(X_A)
    pi_hat = np.zeros(A.shape, dtype=float)
    for tr, te in folds:
        A_tr, X_tr, X_te = A[tr], X_A[tr], X_A[te]
        ctrl = np.sum(A_tr, axis=1) == 0

        def task(j):
            """Inputs of treatment j's fit, or the constant prediction of a degenerate design."""
            elig = mask[tr, j] if mask is not None else (ctrl | (A_tr[:, j] == 1))
            y = A_tr[elig, j]
            if y.size == 0 or not (np.any(y == 1) and np.any(y == 0)):
                raise ValueError(f'Treatment {j} needs at least one eligible control and case in every '
                                 'training fold')
            x = X_tr[elig]
            if np.all(x.max(axis=0) == x.min(axis=0)):
                if class_weight == 'balanced':
                    return None, 0.5
                if isinstance(class_weight, dict):
                    sw = np.asarray([class_weight.get(int(v), 1.0) for v in y])
                    return None, np.average(y, weights=sw)
                return None, np.mean(y)
            return (x, y), None

        # The per-treatment fits are independent (shared controls, own cases): ~1 ms each on the lean
        # path, spread over a few threads (the loss kernels and L-BFGS-B release the GIL).
        n_trt = A.shape[1]
        tasks = [task(j) for j in range(n_trt)]
        todo = [j for j in range(n_trt) if tasks[j][0] is not None]
        fits = [t[1] for t in tasks]
        # (sequential on purpose: the fits are Python-bound -- L-BFGS-B calling back into the loss -- so
        # threads only trade the GIL back and forth; in LFC the whole estimator already runs on a
        # background thread next to the outcome GLM on the device)
        res = [_fit_logistic(lr_kw, tasks[j][0][0], tasks[j][0][1], X_te) for j in todo]
        for j, r in zip(todo, res):
            fits[j] = r
        for j in range(n_trt):
            pi_hat[te, j] = fits[j]
    if clip is not None:
        lo, hi = _validate_clip(clip)
        pi_hat = np.clip(pi_hat, lo, hi)
    return pi_hat


def cross_fitting(Y, A, X, X_A, family='poisson', K=1, glm_alpha=1e-4, ps_model='logistic',
                  ps_class_weight='balanced', Y_hat=None, pi_hat=None, mask=None, ps_clip=(0.01, 0.99),
                  return_raw_pi=False, verbose=False, _glm=None, **kwargs):
    """Nuisance estimation: propensity scores and the outcome model.

    Returns ``(Y_hat, pi_hat[, pi_hat_raw])``.  With ``K == 1`` (in-sample, the default of
    ``LFC``) ``Y_hat`` is a ``YhatFactors`` object; with ``K > 1`` the folds are fitted one by
    one and a dense ``(n, p, a, 2)`` tensor is returned, like the reference.
    """
    kwargs = dict(kwargs)
    if 'class_weight' in kwargs:
        warnings.warn('Passing class_weight through LFC/cross_fitting is deprecated; use ps_class_weight instead.',
                      FutureWarning, stacklevel=2)
        ps_class_weight = kwargs.pop('class_weight')
    params_glm = _filter_params(_glm_mod.fit_glm, {**kwargs, 'verbose': verbose})
    params_glm.pop('_glm', None)
    params_glm.pop('_lazy', None)
    A = np.asarray(A, dtype=float)
    X = np.asarray(X, dtype=float)
    n = X.shape[0]
    if ps_clip is not None and (len(ps_clip) != 2 or not 0 <= ps_clip[0] < ps_clip[1] <= 1):
        raise ValueError('ps_clip must be None or a pair 0 <= lower < upper <= 1')
    ps_future = None
    if pi_hat is None:
        ps_kwargs = {k: v for k, v in kwargs.items() if k != 'random_state'}
        ps_kwargs = _filter_params(LogisticRegression().get_params(), ps_kwargs)

        def _fit_ps():
            return estimate_propensity_scores(A, X_A, K=K, ps_model=ps_model, mask=mask,
                                              random_state=kwargs.get('random_state', 0), verbose=verbose,
                                              class_weight=ps_class_weight, **ps_kwargs)
        if Y_hat is None and K == 1:
            # the host-side logistic fits (sklearn) overlap with the outcome GLM running on the device
            ps_future = _PS_POOL.submit(_fit_ps)
        else:
            with phase('lfc_propensity'):
                pi_hat_raw = _fit_ps()
    else:
        pi_hat_raw = np.asarray(pi_hat, dtype=float).reshape(A.shape)
    if Y_hat is None:
        d = X.shape[1]
        if K > 1:
            Yn = np.asarray(Y, dtype=float)
            if K >= n:
                folds = [([i for i in range(n) if i != j], [j]) for j in range(n)]
            else:
                folds = KFold(n_splits=int(K), random_state=0, shuffle=True).split(X)
            Y_hat = np.zeros((n, Yn.shape[1], A.shape[1], 2))
            for tr, te in folds:
                pg = dict(params_glm)
                if isinstance(pg.get('offset'), np.ndarray):
                    pg['offset'] = pg['offset'][tr]
                B, _, _, off_tr, _ = _glm_mod.fit_glm_auto(Yn[tr], X[tr], A[tr], family=family, alpha=glm_alpha, _lazy=True,
                                                      **pg)
                if not _glm_mod._USE_FAST_BACKEND:
                    # backend="original" only -- the reference's statsmodels route, kept for identical
                    # results: ``fit_glm(..., impute=X_test)`` rebuilds its prediction design from the
                    # TRAINING rows (``X_test = np.c_[X, zeros]`` overwrites the ``impute`` array,
                    # causarray/gcate_glm.py:160-165) and predicts with the training offsets, so the
                    # imputations of the training cells are what ``cross_fitting`` writes into the test
                    # rows (DR_estimation.py:612-613) -- which only broadcasts when both folds have the
                    # same size (K = 2, even n).
                    if len(tr) != len(te):
                        raise ValueError('could not broadcast input array from shape (%d,%d,%d) into shape (%d,%d,%d)'
                                         % (len(tr), Yn.shape[1], A.shape[1], len(te), Yn.shape[1], A.shape[1]))
                    Y_hat[te] = YhatFactors(X[tr], B[:, :d], B[:, d:], off_tr, cap=np.inf).materialize()
                else:
                    # default route (the reference's batched backend, nb_glm_fast.py:273-297): the model
                    # fitted on the training fold predicts the HELD-OUT rows with their own offsets
                    off_all = params_glm.get('offset')
                    if isinstance(off_all, np.ndarray):
                        off_te = off_all[te]
                    elif off_all is True:
                        raise ValueError('cross-fitting needs the offset as an array (offset=True is fold dependent)')
                    else:
                        off_te = None
                    Y_hat[te] = YhatFactors(X[te], B[:, :d], B[:, d:], off_te, cap=np.inf).materialize()
            Y_hat = np.clip(Y_hat, None, 1e5)
        else:
            with phase('lfc_outcome_glm'):
                B, fit, disp, offsets, _ = _glm_mod.fit_glm_auto(Y, X, A, family=family, alpha=glm_alpha, _glm=_glm,
                                                            _lazy=True, **params_glm)
            Y_hat = YhatFactors(X, np.ascontiguousarray(B[:, :d]), np.ascontiguousarray(B[:, d:]), offsets)
            Y_hat.disp_glm = disp
            Y_hat.glm = fit.dev
    elif not isinstance(Y_hat, YhatFactors):
        Y_hat = np.clip(Y_hat, None, 1e5)
    if ps_future is not None:
        with phase('lfc_propensity_wait'):
            pi_hat_raw = ps_future.result()
    pi_hat = pi_hat_raw.copy() if ps_clip is None else np.clip(pi_hat_raw, ps_clip[0], ps_clip[1])
    if return_raw_pi:
        return Y_hat, pi_hat, pi_hat_raw
    return Y_hat, pi_hat


def _repeated_labels(labels, idx):
    """Column ``labels[idx]`` for the result frame.  For string labels the (short) label array is
    converted to pandas' string dtype once and then gathered -- building the 120 000-row column from
    Python objects and letting the DataFrame constructor infer its dtype cost 20 ms per batch."""
    labels = np.asarray(labels)
    if labels.dtype.kind in 'OU' and all(isinstance(v, str) for v in labels.tolist()):
        try:
            return pd.array(labels, dtype='str').take(idx)
        except (TypeError, ValueError):      # pandas without the 'str' dtype alias
            pass
    return labels[idx]


def _add_log2fc_columns(df_res):
    """(Re)compute ``log2fc`` / ``log2fc_se`` right after ``std`` as exact rescalings of ``tau`` / ``std``."""
    ln2 = np.log(2.0)
    l2, l2se = df_res['tau'].to_numpy() / ln2, df_res['std'].to_numpy() / ln2
    for name in ('log2fc', 'log2fc_se'):
        if name in df_res.columns:
            df_res.pop(name)
    at = df_res.columns.get_loc('std') + 1
    df_res.insert(at, 'log2fc', l2)
    df_res.insert(at + 1, 'log2fc_se', l2se)
    return df_res


def compute_causal_estimand(estimand, Y, W, A, W_A=None, family='nb', offset=False, Y_hat=None, pi_hat=None,
                            mask=None, fdx=False, fdx_B=1000, fdx_alpha=0.05, fdx_c=0.1, verbose=False,
                            random_state=0, backend: str = "auto", K=1, ps_clip=(0.01, 0.99),
                            ps_class_weight='balanced', _glm=None, _trusted=False, **kwargs):
    """AIPW estimation of a causal estimand for every (treatment, gene) pair.

    ``estimand`` is either the string ``'LFC'`` together with ``_lfc_opts`` in ``kwargs`` (the
    fused device path used by :func:`LFC`) or a callable ``(etas, A) -> (eta_est, tau, var[, df,
    info])`` as in the reference, which needs the dense pseudo-outcomes and is meant for small
    problems.
    """
    reset_random_seeds(random_state)
    if 'clip_pseudo_outcomes' in kwargs:
        raise TypeError('clip_pseudo_outcomes has been removed; AIPW pseudo-outcomes are always left unclipped')
    lfc_opts = kwargs.pop('_lfc_opts', None)
    ctx = _glm_mod._backend_override(backend) if backend != "auto" else contextlib.nullcontext()
    gene_slice = kwargs.pop('_gene_slice', None)
    if isinstance(Y, pd.DataFrame):
        gene_names = Y.columns
        Y = Y.values
    else:
        gene_names = range(Y.shape[1])
    if hasattr(Y, "tocsr"):
        # sparse counts: with the device copy at hand (_glm) and in-sample nuisance fits nothing on the
        # host needs the dense form; otherwise densify like the reference does
        if _glm is None or K != 1 or Y_hat is not None:
            Y = np.asarray(Y.toarray(), dtype=float)
    elif isinstance(Y, RowsView) and _glm is not None and K == 1 and Y_hat is None:
        pass      # rows of a larger matrix whose device copy is at hand: only the shape is needed here
    else:
        Y = np.asarray(Y)
        # (integer / float32 counts whose device copy is at hand need no float64 copy on the host)
        if _glm is None or K != 1 or Y_hat is not None or Y.dtype.kind not in 'iuf':
            Y = Y.astype(float, copy=False)
    n, p = Y.shape
    # (_trusted: the counts were already validated on the device when the workspace was uploaded)
    if not _trusted and not np.all(np.isfinite(Y.data if hasattr(Y, "tocsr") else Y)):
        raise ValueError('Y must contain only finite values')
    if len(A.shape) == 1:
        A = A.reshape(-1, 1)
    if isinstance(A, pd.DataFrame):
        trt_names = A.columns
        A = A.values
    else:
        trt_names = range(A.shape[1])
    A = np.asarray(A, dtype=float)
    if isinstance(W, pd.DataFrame):
        W = W.values
    if W_A is None:
        W_A = W
    elif isinstance(W_A, pd.DataFrame):
        W_A = W_A.values
    if mask is not None:
        mask = np.array(mask).astype(bool)
        if len(mask.shape) == 1:
            mask = mask.reshape(-1, 1)
        if mask.shape != A.shape:
            raise ValueError('Mask must have the same shape as the treatment matrix')
    kwargs = {k: v for k, v in kwargs.items()
              if k not in ['kwargs_ls_1', 'kwargs_ls_2', 'kwargs_es_1', 'kwargs_es_2', 'c1', 'num_d']}
    if offset is not None and offset is not False:
        if type(offset) == bool and offset is True:
            size_factors = comp_size_factor(Y, **_filter_params(comp_size_factor, kwargs))
            offset = np.log(size_factors)
        else:
            size_factors = np.exp(offset)
    else:
        offset = None
        size_factors = np.ones(n)
    size_factors = np.asarray(size_factors, dtype=float)
    if size_factors.shape != (n,) or not np.all(np.isfinite(size_factors)) or np.any(size_factors <= 0):
        raise ValueError('Size factors must be finite, positive, and have length n')
    if _glm is None and Y_hat is None:
        _glm = GlmDevice(n, p)
        _glm.load_counts(Y)
    with ctx:
        Y_hat, pi_hat, pi_hat_raw = cross_fitting(
            Y, A, W, W_A, family=family, K=K, offset=offset, Y_hat=Y_hat, pi_hat=pi_hat, mask=mask,
            ps_clip=ps_clip, ps_class_weight=ps_class_weight, return_raw_pi=True, random_state=random_state,
            verbose=verbose, _glm=_glm, **kwargs)
    pi_hat = pi_hat.reshape(*A.shape)
    dense = not isinstance(Y_hat, YhatFactors)
    if dense and not np.all(np.isfinite(Y_hat)):
        raise ValueError('Outcome-model predictions must contain only finite values')
    if not dense and not (np.all(np.isfinite(Y_hat.Bcov)) and np.all(np.isfinite(Y_hat.Btrt))):
        raise ValueError('Outcome-model predictions must contain only finite values')
    if not np.all(np.isfinite(pi_hat)) or np.any((pi_hat <= 0) | (pi_hat >= 1)):
        raise ValueError('Propensity scores used by AIPW must be finite and strictly between 0 and 1; '
                         'pass ps_clip bounds or provide valid pi_hat values')

    per_trt = []
    if lfc_opts is not None:
        cm = None if mask is None else mask.astype(np.uint8)
        if dense:
            out = aipw_lfc_yhat(Y, np.asarray(Y_hat, dtype=float), size_factors, A, pi_hat, cell_mask=cm, **lfc_opts)
        else:
            with phase('lfc_aipw'):
                out = aipw_lfc(Y, Y_hat.W, Y_hat.Bcov, Y_hat.Btrt, Y_hat.offset, size_factors, A, pi_hat,
                               cell_mask=cm, glm=getattr(Y_hat, 'glm', None) or _glm, **lfc_opts)
        if gene_slice is not None:     # gene shards: per-gene results of all ranks, in gene order
            from .parallel import allgather_concat
            out = {k_: allgather_concat(v_, axis=1) for k_, v_ in out.items()}
            p = out['tau'].shape[1]
            gene_names = gene_slice[2]
        ctrl = np.sum(A, axis=1) == 0.
        for j in range(A.shape[1]):
            cells = mask[:, j] if mask is not None else (ctrl | (A[:, j] == 1.))
            n0, n1 = int(np.sum(A[cells, j] == 0)), int(np.sum(A[cells, j] == 1))
            if lfc_opts['usevar'] == 'unequal' and (n0 < 2 or n1 < 2):
                warnings.warn(f"Welch variance requires at least 2 cells per arm; got n_0={n0}, n_1={n1} for this "
                              "perturbation. Per-gene variance and df will be NaN; results for these genes are "
                              "silently dropped by downstream BH correction. Pass ``usevar='pooled'`` if you need "
                              "finite estimates in this regime.", RuntimeWarning, stacklevel=3)
            df_eff = out['df'][j] if lfc_opts['usevar'] == 'unequal' else None
            info = {'mean_control': out['mean0'][j], 'mean_treated': out['mean1'][j],
                    'estimable': out['estimable'][j]}
            if df_eff is not None:
                filt = np.isinf(out['var'][j])
                silent = int(np.sum(np.isnan(df_eff) & ~filt))
                if silent > 0:
                    warnings.warn(f"{silent} gene(s) produced NaN Welch df despite passing the low-expression "
                                  "filter; these will be reported as NaN p-values in the result DataFrame.",
                                  RuntimeWarning, stacklevel=3)
            per_trt.append((None, out['tau'][j].copy(), out['var'][j].copy(), df_eff, info))
    else:
        mu = Y_hat if dense else Y_hat.materialize()
        _, etas = AIPW_mean(Y, np.stack([1 - A, A], axis=-1), mu, np.stack([1 - pi_hat, pi_hat], axis=-1))
        etas /= size_factors[:, None, None, None]
        ctrl = np.sum(A, axis=1) == 0.
        for j in range(A.shape[1]):
            cells = mask[:, j] if mask is not None else (ctrl | (A[:, j] == 1.))
            ret = estimand(etas[cells, :, j], A[cells, j], **kwargs)
            per_trt.append((ret[0], ret[1], ret[2], ret[3] if len(ret) > 3 else None,
                            ret[4] if len(ret) > 4 else None))

    _t_inf = phase('lfc_inference_frames')
    _t_inf.__enter__()
    n_trt = A.shape[1]
    if lfc_opts is not None and not fdx:
        # fused path: all treatments at once (one vectorised tail-probability call, one frame)
        tau_all = np.stack([q[1] for q in per_trt])
        var_all = np.stack([q[2] for q in per_trt])
        df_all = np.stack([q[3] for q in per_trt]) if per_trt[0][3] is not None else None
        with np.errstate(invalid='ignore', divide='ignore'):
            std_all = np.sqrt(var_all)
            t_all = tau_all / std_all
        t_all[np.isinf(std_all)] = np.nan
        pv, qv, pva, qva = bh_correction_device(t_all, df=df_all)
        gn = np.asarray(list(gene_names))
        cols = {'gene_names': _repeated_labels(gn, np.tile(np.arange(p), n_trt)), 'tau': tau_all.ravel(),
                'std': std_all.ravel(),
                'stat': t_all.ravel(), 'rej': np.zeros(n_trt * p), 'pvalue': pv.ravel(), 'padj': qv.ravel(),
                'pvalue_emp_null_adj': pva.ravel(), 'padj_emp_null_adj': qva.ravel(),
                'mean_control': np.concatenate([q[4]['mean_control'] for q in per_trt]),
                'mean_treated': np.concatenate([q[4]['mean_treated'] for q in per_trt]),
                'estimable': np.concatenate([q[4]['estimable'] for q in per_trt])}
        if n_trt > 1:
            cols['trt'] = _repeated_labels(np.asarray(list(trt_names), dtype=object), np.repeat(np.arange(n_trt), p))
        df_res = pd.DataFrame(cols)
        for j in range(n_trt):
            n_bad = int(np.sum(~np.asarray(per_trt[j][4]['estimable'], dtype=bool)))
            if n_bad:
                trt = trt_names[j] if n_trt > 1 else 0
                warnings.warn(f'{n_bad} gene(s) for treatment {trt!r} have nonfinite or nonpositive AIPW '
                              'counterfactual means and are reported as non-estimable.', RuntimeWarning,
                              stacklevel=2)
    else:
        frames = []
        for j, (eta_est, tau_est, var_est, df_est, info) in enumerate(per_trt):
            with np.errstate(invalid='ignore', divide='ignore'):
                std_est = np.sqrt(var_est)
                tvals = tau_est / std_est
            if fdx:
                if eta_est is None:
                    raise NotImplementedError('fdx=True needs the dense influence values; pass a callable estimand')
                V = fdx_control(tau_est, tvals, eta_est, std_est, fdx, fdx_B, fdx_alpha, fdx_c)
            else:
                V = np.zeros(tvals.shape[0])
            tvals[np.isinf(std_est)] = np.nan
            pv, qv, pva, qva = bh_correction(tvals, df=df_est)
            df_j = pd.DataFrame({'gene_names': gene_names, 'tau': tau_est, 'std': std_est, 'stat': tvals, 'rej': V,
                                 'pvalue': pv, 'padj': qv, 'pvalue_emp_null_adj': pva, 'padj_emp_null_adj': qva})
            if info is not None:
                for name, values in info.items():
                    df_j[name] = values
                n_bad = int(np.sum(~np.asarray(info['estimable'], dtype=bool)))
                if n_bad:
                    trt = trt_names[j] if n_trt > 1 else 0
                    warnings.warn(f'{n_bad} gene(s) for treatment {trt!r} have nonfinite or nonpositive AIPW '
                                  'counterfactual means and are reported as non-estimable.', RuntimeWarning,
                                  stacklevel=2)
            if n_trt > 1:
                df_j['trt'] = trt_names[j]
            frames.append(df_j)
        df_res = pd.concat(frames, axis=0).reset_index(drop=True)
    _t_inf.__exit__(None, None, None)
    estimation = {**{'pi_hat': pi_hat, 'pi_hat_raw': pi_hat_raw, 'Y_hat': Y_hat, 'offset': offset,
                     'size_factors': size_factors, 'ps_class_weight': ps_class_weight}, **kwargs}
    return df_res, estimation


def LFC(Y, W, A, W_A=None, family='nb', offset=False, Y_hat=None, pi_hat=None, cross_est=False, K=None,
        mask=None, usevar: Literal['unequal', 'pooled'] = 'unequal', thres_min=1e-2, thres_diff=1e-2,
        eps_var=1e-4, fdx=False, fdx_alpha=0.05, fdx_c=0.1, verbose=False, backend: str = "auto",
        ps_clip=(0.01, 0.99), ps_class_weight='balanced', **kwargs):
    """Doubly-robust log-fold changes ``log E[Y(1)] / E[Y(0)]`` with Welch (``'unequal'``) or pooled
    variance, low-expression filter, t / BH inference.

    Returns ``(df, estimation)``; ``df`` columns in order: ``gene_names, tau, std, log2fc,
    log2fc_se, stat, rej, pvalue, padj, pvalue_emp_null_adj, padj_emp_null_adj, mean_control,
    mean_treated, estimable[, trt]``.
    """
    if usevar not in ('pooled', 'unequal'):
        raise ValueError('usevar must be either "pooled" or "unequal"')
    if K is None:
        K = 2 if cross_est else 1
    if kwargs.pop('gene_shard', False):
        # genes split over the ranks (every rank makes the same call with the full inputs): the outcome
        # GLM and the AIPW reduction run on this rank's genes, the per-gene results are gathered and the
        # per-treatment BH runs on all genes; every rank returns the full frame
        from .parallel import dist_info, gene_bounds
        rank, world = dist_info()
        if world > 1:
            names_full = list(Y.columns) if isinstance(Y, pd.DataFrame) else None
            Yf = Y.values if isinstance(Y, pd.DataFrame) else (Y.toarray() if hasattr(Y, 'tocsr') else np.asarray(Y))
            p_total = Yf.shape[1]
            if fdx or K != 1 or (Y_hat is not None and not isinstance(Y_hat, YhatFactors)):
                raise NotImplementedError('gene_shard=True supports the in-sample fused path only '
                                          '(no fdx, K = 1, no dense Y_hat)')
            if type(offset) == bool and offset is True:
                # median-of-ratios size factors are a per-cell median over ALL genes: compute them from
                # the full matrix before the columns are sliced (every rank gets the same offsets)
                offset = np.log(comp_size_factor(Yf, **_filter_params(comp_size_factor, kwargs)))
            lo, hi = gene_bounds(p_total, world)[rank]
            if kwargs.get('_glm') is None or kwargs['_glm'].p != p_total:
                Y = np.ascontiguousarray(Yf[:, lo:hi])
            if kwargs.get('disp_glm') is not None and np.size(kwargs['disp_glm']) == p_total:
                kwargs['disp_glm'] = np.ravel(kwargs['disp_glm'])[lo:hi]
            kwargs['_gene_slice'] = (lo, hi, names_full if names_full is not None else range(p_total))
    opts = dict(usevar=usevar, thres_min=thres_min, thres_diff=thres_diff, eps_var=eps_var)
    if fdx:
        # FDX needs the per-cell influence values: dense reference-style estimand (small problems)
        estimand = _dense_lfc_estimand(**opts)
        df_res, estimation = compute_causal_estimand(
            estimand, Y, W, A, W_A, family, offset, Y_hat=Y_hat, pi_hat=pi_hat, mask=mask, fdx=fdx,
            fdx_alpha=fdx_alpha, fdx_c=fdx_c, verbose=verbose, backend=backend, K=K, ps_clip=ps_clip,
            ps_class_weight=ps_class_weight, **kwargs)
    else:
        df_res, estimation = compute_causal_estimand(
            'LFC', Y, W, A, W_A, family, offset, Y_hat=Y_hat, pi_hat=pi_hat, mask=mask, fdx=False,
            fdx_alpha=fdx_alpha, fdx_c=fdx_c, verbose=verbose, backend=backend, K=K, ps_clip=ps_clip,
            ps_class_weight=ps_class_weight, _lfc_opts=opts, **kwargs)
    return _add_log2fc_columns(df_res), estimation


def _dense_lfc_estimand(usevar, thres_min, thres_diff, eps_var):
    """Host estimand on dense pseudo-outcomes (only used when ``fdx=True``)."""
    def estimand(etas, A, **kw):
        e0, e1 = etas[..., 0], etas[..., 1]
        m0, m1 = np.mean(e0, axis=0, dtype=np.float64), np.mean(e1, axis=0, dtype=np.float64)
        est = np.isfinite(m0) & np.isfinite(m1) & (m0 > 0) & (m1 > 0)
        t0 = np.where(est, np.maximum(m0, thres_diff), 1.0)
        t1 = np.where(est, np.maximum(m1, thres_diff), 1.0)
        with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
            tau = np.log(t1 / t0)
            IF = e1 / t1[None, :] - e0 / t0[None, :]
        df_eff = None
        if usevar == 'pooled':
            var = (np.var(IF, axis=0, ddof=1) + eps_var) / IF.shape[0]
        else:
            n0, n1 = int(np.sum(A == 0)), int(np.sum(A == 1))
            with np.errstate(invalid='ignore', divide='ignore'):
                v0 = (np.var(IF[A == 0], axis=0, ddof=1) + eps_var) / n0
                v1 = (np.var(IF[A == 1], axis=0, ddof=1) + eps_var) / n1
                var = v0 + v1
                df_eff = (v0 + v1) ** 2 / (v0 ** 2 / (n0 - 1) + v1 ** 2 / (n1 - 1))
        drop = ~est | (np.maximum(m0, m1) < thres_min) | (np.abs(m1 - m0) < thres_diff)
        tau[drop] = 0.
        IF[:, drop] = 0.
        var[drop] = np.inf
        if df_eff is not None:
            df_eff[drop] = np.nan
        return IF, tau, var, df_eff, {'mean_control': m0, 'mean_treated': m1, 'estimable': est}
    return estimand


def _cache_put(cache_path, batch_i, df_b, family, a_total):
    with pd.HDFStore(cache_path, mode='a') as store:
        if '/meta' not in store.keys():
            store.put('/meta', pd.DataFrame({'causarray_version': [__version__], 'family': [family],
                                             'a_total': [a_total],
                                             'columns': [','.join(df_b.columns.astype(str))]}), format='fixed')
        store.put(f'batch_{batch_i:04d}', df_b, format='fixed')


def gcate_lfc_batch(Y, X, A, r, W_A=None, batch_size=10, n_batches=None, max_cells=2000, n_ctrl=2000,
                    family='nb', offset=True, warm_start_U=False, cache_path=None, random_state=0, verbose=False,
                    gcate_kwargs=None, lfc_kwargs=None, **kwargs):
    """Batch-wise GCATE + doubly-robust LFC over chunks of perturbations.

    Same call and result as the reference (rows batch-major, then treatment-major, then gene
    order; columns of :func:`LFC` plus ``trt`` and ``batch``; optional resumable HDF5 cache).
    The counts of each batch are uploaded once and stay on the device through the GCATE fit,
    the outcome GLM and the AIPW reduction.
    """
    _t_call = time.perf_counter()
    from .estimators import _dense_counts
    import scipy.sparse as _sp
    gcate_kwargs = {} if gcate_kwargs is None else gcate_kwargs
    lfc_kwargs = {} if lfc_kwargs is None else lfc_kwargs
    gene_names = list(Y.columns) if isinstance(Y, pd.DataFrame) else None
    sparse_in = _sp.issparse(Y)      # kept sparse: one batch of rows is densified at a time
    Y_np = Y.tocsr() if sparse_in else _dense_counts(Y)
    X_np = np.asarray(X)
    W_A_np = np.asarray(W_A) if W_A is not None else None
    if isinstance(A, pd.DataFrame):
        A_np = A.values.astype(float)
        names = list(A.columns)
    else:
        A_np = np.asarray(A, dtype=float)
        names = list(range(A_np.shape[1]))
    cached, skip = {}, set()
    if cache_path is not None:
        meta = None
        try:
            with pd.HDFStore(cache_path, mode='r') as store:
                keys = store.keys()
                if '/meta' in keys:
                    meta = store['/meta']
                for key in keys:
                    if key.startswith('/batch_'):
                        try:
                            idx = int(key.split('_')[1])
                            cached[idx] = key
                            skip.add(idx)
                        except (ValueError, IndexError):
                            pass
        except (FileNotFoundError, OSError):
            pass
        if meta is not None:
            for k_, v_ in {'family': family, 'a_total': int(A_np.shape[1])}.items():
                if k_ in meta.columns and meta.iloc[0][k_] != v_:
                    raise ValueError(f"gcate_lfc_batch cache at {cache_path!r} was written with "
                                     f"{k_}={meta.iloc[0][k_]!r}, but the current call uses {k_}={v_!r}.  Refusing "
                                     "to mix incompatible schemas -- delete the cache file or pass a fresh "
                                     "cache_path to start over.")
    col_of = {name: i for i, name in enumerate(names)}
    new = {}
    if os.environ.get('CA_TIMING'):
        print('[gcate_lfc_batch] set-up %.3f s' % (time.perf_counter() - _t_call), file=sys.stderr)
    # multi-GPU: every rank takes the batches i with i % world == rank (no data-path collective);
    # the frames are gathered on rank 0 at the end (causarray_b200/parallel.py)
    from . import parallel as _par
    import math as _math
    _rank, _world = _par.dist_info()
    if _world > 1:
        if warm_start_U:
            # U of batch i-1 would come from a different rank: the result would depend on WORLD_SIZE
            raise ValueError('warm_start_U=True chains the batches and cannot be combined with batch sharding '
                             'over several ranks; run it on one rank or pass warm_start_U=False')
        _nb = n_batches if n_batches is not None else _math.ceil(A_np.shape[1] / batch_size)
        _nb = max(1, min(_nb, A_np.shape[1]))
        skip = set(skip) | _par.foreign_batches(_nb, _rank, _world)
    # one batch at a time: GCATE fit -> LFC on the same device-resident counts
    from .estimators import iter_gcate_batches
    # Python's cyclic collector is paused for the loop (the reference calls gc.collect() after every
    # batch, causarray/DR_learner.py:846; with the result frames of the finished batches alive a full
    # collection -- explicit or triggered by the allocations of a batch -- costs ~50 ms, a quarter of a
    # batch): device memory is released deterministically by ws.close(), host arrays by reference counts.
    _gc_was_on = gc.isenabled()
    gc.disable()
    try:
        for br in iter_gcate_batches(Y_np, X_np, A, r, batch_size=batch_size, n_batches=n_batches, max_cells=max_cells,
                                     n_ctrl=n_ctrl, family=family, offset=offset, warm_start_U=warm_start_U,
                                     skip_batches=skip, random_state=random_state, verbose=verbose,
                                     **{**kwargs, **gcate_kwargs}):
            if br.get('skipped'):
                continue
            batch_i, cell_idx, res_2 = br['batch_i'], br['cell_idx'], br['res_2']
            ws = res_2.pop('_ws', None)
            U_b = res_2['U'].copy()
            off_b = np.log(res_2['kwargs_glm']['size_factor'])
            # (the counts are on the device already: a sparse slice stays sparse, a dense one stays a view)
            if gene_names is not None and not sparse_in:
                Y_b = pd.DataFrame(Y_np[cell_idx], columns=gene_names)
            else:
                Y_b = RowsView(Y_np, cell_idx) if (ws is not None and not sparse_in) else Y_np[cell_idx]
            X_b = X_np[cell_idx]
            cols = [col_of[name] for name in br['pert_names']]
            A_b = A_np[np.ix_(cell_idx, cols)]
            W_b = np.c_[X_b, U_b]
            W_A_b = np.c_[W_A_np[cell_idx], U_b] if W_A_np is not None else W_b
            df_b, est_b = LFC(Y_b, W_b, pd.DataFrame(A_b, columns=br['pert_names']), W_A_b, family=family,
                              offset=off_b, verbose=verbose, _glm=None if ws is None else ws.glm,
                              _trusted=ws is not None, **{**kwargs, **lfc_kwargs})
            df_b['batch'] = batch_i
            new[batch_i] = df_b
            if cache_path is not None and _world == 1:
                _cache_put(cache_path, batch_i, df_b, family, int(A_np.shape[1]))
            del est_b
            if ws is not None:
                ws.close()
            br['res_1'] = None
            br['res_2'] = None
    finally:
        if _gc_was_on:
            gc.enable()
    _t_loop = time.perf_counter()
    if _world > 1:
        new = _par.gather_frames(new)
        if os.environ.get('CA_TIMING'):
            print('[gcate_lfc_batch] rank %d: batches %.3f s, gather %.3f s' % (_rank, _t_loop - _t_call,
                                                                                  time.perf_counter() - _t_loop), file=sys.stderr)
        if new is None:
            return None   # only rank 0 returns the table
        if cache_path is not None:   # HDF5 is not concurrent-safe: rank 0 writes everything
            for _i in sorted(new):
                _cache_put(cache_path, _i, new[_i], family, int(A_np.shape[1]))
    idx_all = sorted(set(new) | set(cached))
    if cached and cache_path is not None:
        with pd.HDFStore(cache_path, mode='r') as store:
            frames = [new[i] if i in new else store[cached[i]] for i in idx_all]
            result = pd.concat(frames, axis=0).reset_index(drop=True)
    else:
        result = pd.concat([new[i] for i in idx_all], axis=0).reset_index(drop=True)
    result = _add_log2fc_columns(result)
    if os.environ.get('CA_TIMING'):
        print('[gcate_lfc_batch] rank %d: batches done at %.3f s, table at %.3f s' % (_rank, _t_loop - _t_call,
                                                                                      time.perf_counter() - _t_call), file=sys.stderr)
    return result


def LFC_batch(*args, **kwargs):
    """Deprecated alias of :func:`gcate_lfc_batch`."""
    warnings.warn("LFC_batch is deprecated and will be removed in a future release. Use gcate_lfc_batch instead.",
                  DeprecationWarning, stacklevel=2)
    return gcate_lfc_batch(*args, **kwargs)
