"""Per-gene GLM fits on the device -- the S2 seam of SURVEY.md section 8(b).

Mirrors the reference interface of ``causarray/gcate_glm.py`` (``fit_glm``
``:95-272``, ``estimate_disp`` ``:275-308``, ``fit_glm_auto`` ``:365-457``,
``estimate_disp_auto`` ``:460-476``, ``_backend_override`` ``:53-76``) and of
``causarray/nb_glm_fast.py`` (``fit_glm_fast`` ``:164-301``, deviance residuals
``:729-777``).  All genes are fitted at once by ``ca_glm_fit`` (batched IRLS
with FP64 tensor-core Gram contractions, ``csrc/glm.cu``); the arithmetic is
the joint-MLE ("original" / statsmodels) one -- the reference's crispyx "fast"
approximation is an un-vendored third-party algorithm and is not imitated:
``fit_glm_fast`` and ``backend="fast"`` route to the same device fit.

Genes whose unpenalised fit trips the reference's guards (non-finite
coefficients, |beta_cov| > 50, |beta_trt| > 10; ``gcate_glm.py:205``) are refitted
on the device by a lasso-penalised IRLS whose fixed point minimises
``-loglike / n + sum_k alpha_k |beta_k|`` -- the objective of the statsmodels
``fit_regularized(alpha=1e-4 on non-constant columns, L1_wt=1)`` call the
reference falls back to (``:209``; statsmodels stops its coordinate descent
at ``cnvrg_tol=1e-5``, the device solves to 1e-8 in the penalised deviance);
their deviance residuals are zero, as in the reference (``:210``).  The number of such genes is
reported in ``last_fit_info``.
"""
import contextlib
import pprint
import warnings

import numpy as np

from .device import GlmDevice
from .prep import comp_size_factor, _filter_params

_USE_FAST_BACKEND = True
_FAST_ROUTE = False        # set inside _backend_override("fast"): fit_glm_auto takes the two-stage route
_FAST_MIN_P = 50
_FAST_MAX_D = 50
_FAST_MAX_COEF = 1e4
_CRISPYX_AVAILABLE = False   # the device backend replaces crispyx; kept for API compatibility

last_fit_info = {}
COUNTERS = {'irls_gene_iters': 0, 'fits': 0, 'gram_flops': 0.0, 'gram_flops_structured': 0.0}


@contextlib.contextmanager
def _backend_override(backend):
    """Scoped ``"fast"`` / ``"original"`` / ``"auto"`` switch (all run the same device IRLS)."""
    global _USE_FAST_BACKEND, _FAST_ROUTE
    old, old_route = _USE_FAST_BACKEND, _FAST_ROUTE
    if backend == "fast":
        _USE_FAST_BACKEND, _FAST_ROUTE = True, True
    elif backend == "original":
        _USE_FAST_BACKEND, _FAST_ROUTE = False, False
    try:
        yield
    finally:
        _USE_FAST_BACKEND, _FAST_ROUTE = old, old_route


class GlmFit:
    """Result of one batched device fit; heavy (n, p) arrays are pulled lazily."""

    def __init__(self, dev, B, deviance, n_iter, status, refit_mask):
        self.dev, self.B, self.deviance, self.n_iter, self.status = dev, B, deviance, n_iter, status
        self.refit_mask = refit_mask

    def fitted(self):
        return self.dev.fitted()

    def resid(self):
        return self.dev.resid()


def _device_fit(dev, Xf, family, offsets, disp, thres_disp, maxiter, d_guard, alpha=1e-4, verbose=False):
    """Unpenalised IRLS for all genes, then the lasso refit of the genes that trip the reference's
    coefficient guards (``causarray/gcate_glm.py:205-209``: ``fit_regularized(alpha)``, an L1
    penalty ``alpha = 1e-4`` on every non-constant column)."""
    n = Xf.shape[0]
    fam = 'poisson' if family == 'poisson' else 'nb'
    nu = None if fam == 'poisson' else np.asarray(disp, dtype=np.float64)
    B, dv, n_iter, status = dev.fit(fam, Xf, offsets, nu, thres_disp=thres_disp, maxiter=maxiter, tol=1e-8,
                                    d_guard=d_guard)
    bad = status != 0
    if bad.any():
        const = np.all(Xf == Xf[0, :], axis=0)
        l1 = np.where(const, 0.0, alpha * n)
        B2, dv2, it2, st2 = dev.refit_l1(fam, Xf, l1, bad, offset=offsets, nu=nu, thres_disp=thres_disp, maxiter=200,
                                         tol=1e-8)
        still = bad & (st2 != 0)
        B = B2
        dv = dv2
        if still.any():
            B[still] = 0.0
            if verbose:
                pprint.pprint('Fitting GLM for {} columns does not converge.'.format(int(still.sum())))
    COUNTERS['irls_gene_iters'] += int(n_iter.sum())
    COUNTERS['fits'] += 1
    _kx = Xf.shape[1]
    # algorithmic FLOPs of the IRLS passes: 2 n [kx(kx+1)/2 + 2 kx] per gene and iteration
    COUNTERS['gram_flops'] += 2.0 * n * float(n_iter.sum()) * (_kx * (_kx + 1) / 2 + 2 * _kx)
    # the same with the one-hot block of the design (at most one non-zero 0/1 entry per cell) taken out:
    # only the dense columns form pair products (csrc/glm_irls_impl.cuh)
    _kd = _dense_width(Xf)
    COUNTERS['gram_flops_structured'] += 2.0 * n * float(n_iter.sum()) * (_kd * (_kd + 1) / 2 + 2 * _kd + 65)
    last_fit_info.update(n_genes=int(B.shape[0]), n_refit=int(bad.sum()), max_iter=int(n_iter.max(initial=0)),
                         irls_gene_iters=int(n_iter.sum()))
    return GlmFit(dev, B, dv, n_iter, status, bad)


_DENSE_WIDTH_CACHE = {}


def _dense_width(Xf):
    """Number of design columns outside the largest contiguous one-hot block (bench accounting only)."""
    key = (Xf.shape, float(Xf[:, -1].sum()), float(np.abs(Xf[0]).sum()))
    if key not in _DENSE_WIDTH_CACHE:
        binary = np.all((Xf == 0) | (Xf == 1), axis=0) & ~np.all(Xf == 1, axis=0)
        best, lo = 0, 0
        kx = Xf.shape[1]
        while lo < kx:
            hi = lo
            while hi < kx and binary[hi] and np.all(Xf[:, lo:hi + 1].sum(axis=1) <= 1):
                hi += 1
            best = max(best, hi - lo)
            lo = max(hi, lo + 1)
        _DENSE_WIDTH_CACHE[key] = kx - (best if best >= 2 else 0)
    return _DENSE_WIDTH_CACHE[key]


def _as_device(Y, _glm):
    if _glm is not None:
        return _glm
    Y = np.asarray(Y, dtype=np.float64)
    dev = GlmDevice(Y.shape[0], Y.shape[1])
    dev.load_counts(Y)
    return dev


def fit_glm(Y, X, A=None, family='gaussian', disp_family='poisson',
            disp_glm=None, impute=False, offset=None, offset_test=None, shrinkage=False,
            alpha=1e-4, maxiter=1000, thres_disp=100., n_jobs=-3, random_state=0, verbose=False,
            mem_limit_gb=None, _glm=None, _lazy=False, **kwargs):
    """Fit one GLM per column of ``Y`` on covariates ``X`` and treatments ``A``.

    Returns ``(B, Yhat, disp_glm, offsets, resid_deviance)`` like the reference:
    ``B (p, d[+a])``; ``Yhat`` the fitted means ``(n, p)`` or, with ``impute``,
    the counterfactual pair ``(Yhat_0, Yhat_1)`` each ``(n, p, a)``.
    """
    np.random.seed(random_state)
    if family not in ['gaussian', 'poisson', 'nb']:
        raise ValueError('Family not recognized')
    if family == 'gaussian':
        raise NotImplementedError('the gaussian family is outside the accelerated GCATE / LFC path')
    X = np.asarray(X, dtype=np.float64)
    d = X.shape[1]
    if A is None:
        a = 1
        assert impute is False
        Xf = X
    else:
        A = np.asarray(A, dtype=np.float64)
        if A.ndim == 1:
            A = A[:, None]
        a = A.shape[1]
        Xf = np.c_[X, A]
    if offset is not None and offset is not False:
        if type(offset) == bool and offset is True:
            offsets = np.log(comp_size_factor(Y, **_filter_params(comp_size_factor, kwargs)))
        else:
            offsets = np.asarray(offset, dtype=np.float64)
    else:
        offsets = None
    dev = _as_device(Y, _glm)
    if family == 'nb' and disp_glm is None:
        disp_glm = estimate_disp(Y, Xf, offset=offsets, disp_family=disp_family, maxiter=1000, verbose=verbose,
                                 _glm=dev, **kwargs)
    if verbose:
        pprint.pprint('Fitting {} GLM{}...'.format(family, '' if offsets is None else ' with offset'))
    fit = _device_fit(dev, Xf, family, offsets, disp_glm, thres_disp, maxiter, d, alpha=alpha, verbose=verbose)
    B = fit.B
    if _lazy:
        return B, fit, disp_glm, offsets, None
    resid = fit.resid()
    if fit.refit_mask.any():
        resid[:, fit.refit_mask] = 0.0
    if impute is not False and A is not None:
        eta0 = X @ B[:, :d].T
        if offsets is not None:
            eta0 = eta0 + offsets[:, None]
        dt = np.float64
        if mem_limit_gb is not None and eta0.size * a * 2 * 8 / 1e9 > mem_limit_gb:
            warnings.warn("Imputation arrays exceed mem_limit_gb; downcasting to float32.", ResourceWarning,
                          stacklevel=2)
            dt = np.float32
        with np.errstate(over='ignore'):
            Yhat_0 = np.repeat(np.exp(eta0)[:, :, None], a, axis=2).astype(dt)
            Yhat_1 = np.exp(eta0[:, :, None] + B[:, d:][None, :, :]).astype(dt)
        Yhat = (Yhat_0, Yhat_1)
    else:
        Yhat = fit.fitted()
    return B, Yhat, disp_glm, offsets, resid


def estimate_disp(Y, X=None, A=None, Y_hat=None, disp_family='gaussian', offset=None, verbose=False, _glm=None,
                  **kwargs):
    """Method-of-moments NB size parameter (``1 / clip(phi, 0.01, 100)``) per gene."""
    if offset is not None:
        if type(offset) == bool and offset is True:
            offsets = np.log(comp_size_factor(Y, **_filter_params(comp_size_factor, kwargs)))
        else:
            offsets = np.asarray(offset, dtype=np.float64)
    else:
        offsets = None
    if Y_hat is None:
        if A is not None:
            X = np.c_[X, A]
        if disp_family != 'poisson':
            raise NotImplementedError("only disp_family='poisson' runs on the device")
        dev = _as_device(Y, _glm)
        X = np.asarray(X, dtype=np.float64)
        _device_fit(dev, X, 'poisson', offsets, None, 100., int(kwargs.get('maxiter', 1000)), X.shape[1])
        return dev.disp_moments(offsets)
    # explicit fitted values (already on the normalised scale): small host formula
    Y = np.asarray(Y, dtype=np.float64)
    sf = 1. if offsets is None else np.exp(offsets)[:, None]
    Yn = Y / sf
    Yh = np.clip(Y_hat, 0., np.max(Yn, axis=0))
    with np.errstate(divide='ignore', invalid='ignore'):
        phi = np.mean((Yn - Yh) ** 2 - Yh, axis=0) / np.mean(Yh ** 2, axis=0)
        disp = 1. / np.clip(phi, 0.01, 100.)
    disp[np.isnan(disp)] = 1.
    return disp


def fit_glm_auto(Y, X, A=None, family='gaussian', disp_family='poisson',
                 disp_glm=None, impute=False, offset=None, offset_test=None, shrinkage=False,
                 alpha=1e-4, maxiter=1000, thres_disp=100., n_jobs=-3, random_state=0,
                 verbose=False, mem_limit_gb=None, **kwargs):
    """Backend router of the reference (``causarray/gcate_glm.py:365-457``): ``backend="fast"`` (the
    ``_backend_override("fast")`` scope) with at least ``_FAST_MIN_P`` genes and at most ``_FAST_MAX_D``
    design columns takes the two-stage batched route (:func:`fit_glm_fast`) and falls back to the joint
    fit when a coefficient exceeds ``_FAST_MAX_COEF`` in magnitude or is not finite, as the reference
    does; everything else -- including ``"auto"``, which is what the reference runs where crispyx is not
    installed -- is the joint per-gene MLE (:func:`fit_glm`)."""
    if _FAST_ROUTE and family in ('poisson', 'nb') and Y.shape[1] >= _FAST_MIN_P \
            and np.asarray(X).shape[1] + (0 if A is None else np.atleast_2d(np.asarray(A).T).shape[0]) <= _FAST_MAX_D \
            and not shrinkage:
        try:
            out = fit_glm_fast(Y, X, A=A, family=family, disp_family=disp_family, disp_glm=disp_glm, impute=impute,
                               offset=offset, offset_test=offset_test, alpha=alpha, maxiter=maxiter,
                               thres_disp=thres_disp, random_state=random_state, verbose=verbose,
                               mem_limit_gb=mem_limit_gb, **kwargs)
            if np.all(np.isfinite(out[0])) and np.max(np.abs(out[0]), initial=0.0) <= _FAST_MAX_COEF:
                return out
            warnings.warn('fast GLM route produced non-finite or very large coefficients; using the joint fit',
                          RuntimeWarning, stacklevel=2)
        except NotImplementedError:
            pass
    return fit_glm(Y, X, A=A, family=family, disp_family=disp_family, disp_glm=disp_glm, impute=impute,
                   offset=offset, offset_test=offset_test, shrinkage=shrinkage, alpha=alpha, maxiter=maxiter,
                   thres_disp=thres_disp, n_jobs=n_jobs, random_state=random_state, verbose=verbose,
                   mem_limit_gb=mem_limit_gb, **kwargs)


_N_STAGE1 = 3000        # cells of the global covariate fit (causarray/nb_glm_fast.py:418)


def fit_glm_fast(Y, X, A=None, family='gaussian', disp_family='poisson', disp_glm=None, impute=False, offset=None,
                 offset_test=None, shrinkage=False, alpha=1e-4, maxiter=1000, thres_disp=100., n_jobs=-3,
                 random_state=0, verbose=False, mem_limit_gb=None, _lazy=False, **kwargs):
    """The reference's batched "fast" route (``causarray/nb_glm_fast.py:164-594``), structure for structure,
    on the device IRLS.

    * no treatments, or a single one: one joint fit capped at ``min(maxiter, 50)`` iterations
      (``_fit_glm_fast_single``, ``:304-370``);
    * several treatments (``_fit_glm_fast_per_perturbation``, ``:373-594``): **stage 1** -- covariate-only
      model on at most 3 000 cells (``np.random.default_rng(random_state).choice``, sorted), a Poisson
      fit capped at 5 (NB) / 10 (Poisson) iterations, the method-of-moments dispersion from its fitted
      means (clipped to alpha in [1e-8, 100]) and, for NB, an NB fit capped at 5 iterations; **stage 2** --
      for every perturbation the 1-D treatment effect against the covariate offset ``X B_cov^T`` with that
      dispersion held fixed, at most 50 iterations: with one-hot indicators the perturbations decouple,
      so all of them are one launch of ``ca_glm_fit_treatment`` (control cells do not enter a
      treatment's score equation, exactly as in the reference's ctrl + pert_k sub-fits).

    The arithmetic inside the reference's stages is crispyx's ``NBGLMBatchFitter`` (un-vendored, absent
    here): PARITY UNPINNED for it -- this function reproduces the documented structure (subsample, caps,
    offsets, fixed dispersion, cell-weighted dispersion average) with the statsmodels-form IRLS of
    ``fit_glm``; ``oracle/glm_fast.py`` restates it for the tests.  Returns like ``fit_glm``.
    """
    np.random.seed(random_state)
    if family not in ['gaussian', 'poisson', 'nb']:
        raise ValueError('Family not recognized')
    if family == 'gaussian':
        raise NotImplementedError('the gaussian family is outside the accelerated GCATE / LFC path')
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    if A is not None:
        A = np.asarray(A, dtype=np.float64)
        if A.ndim == 1:
            A = A[:, None]
    if offset is not None and offset is not False:
        if type(offset) == bool and offset is True:
            offsets = np.log(comp_size_factor(Y, **_filter_params(comp_size_factor, kwargs)))
        else:
            offsets = np.asarray(offset, dtype=np.float64).ravel()
    else:
        offsets = None
    one_hot = A is not None and A.shape[1] > 1 and np.all((A == 0) | (A == 1)) and np.all(A.sum(axis=1) <= 1)
    if A is None or A.shape[1] == 1 or not one_hot:
        # single joint batch fit (iteration cap of the reference's fast path)
        return fit_glm(Y, X, A=A, family=family, disp_family=disp_family, disp_glm=disp_glm, impute=impute,
                       offset=offsets, shrinkage=shrinkage, alpha=alpha, maxiter=min(int(maxiter), 50),
                       thres_disp=thres_disp, random_state=random_state, verbose=verbose,
                       mem_limit_gb=mem_limit_gb, _lazy=_lazy, **kwargs)
    a = A.shape[1]
    Yd = np.asarray(Y)
    p = Yd.shape[1]
    # ---- stage 1: global covariate model on a cell subsample ----
    if n <= _N_STAGE1:
        s1 = np.arange(n)
    else:
        s1 = np.sort(np.random.default_rng(random_state).choice(n, size=_N_STAGE1, replace=False))
    dev1 = GlmDevice(len(s1), p)
    dev1.load_counts(np.ascontiguousarray(Yd[s1], dtype=np.float64))
    off1 = None if offsets is None else offsets[s1]
    cap = 5 if family == 'nb' else 10
    Bp, _, it_p, _ = dev1.fit('poisson', X[s1], off1, None, thres_disp=thres_disp, maxiter=min(int(maxiter), cap),
                              tol=1e-8, d_guard=-1)
    n_it = int(it_p.sum())
    if family == 'nb':
        # (the dispersion is a gene-level property estimated on the covariate-only design, :239-246; a
        # caller-supplied disp_glm is used for the NB weights of stage 1 as the reference passes it on)
        disp1 = dev1.disp_moments(off1)
        alpha_g = np.clip(1.0 / disp1, 1e-8, 100.0)
        alpha_g[~np.isfinite(alpha_g)] = 1.0
        nu_g = 1.0 / alpha_g
        B_cov, _, it_n, _ = dev1.fit('nb', X[s1], off1, nu_g, thres_disp=np.inf, maxiter=min(int(maxiter), 5), tol=1e-8,
                                     d_guard=-1)
        n_it += int(it_n.sum())
    else:
        B_cov, nu_g = Bp, None
    dev1.close()
    # ---- stage 2: per-perturbation treatment effects against the covariate offset, dispersion fixed ----
    dev = _as_device(Y, kwargs.get('_glm'))
    Xf = np.c_[X, A]
    B0 = np.c_[B_cov, np.zeros((p, a))]
    B, dv, it2, st2 = dev.fit_treatment('poisson' if family == 'poisson' else 'nb', Xf, B0, offsets, nu_g,
                                        thres_disp=np.inf, maxiter=min(int(maxiter), 50), tol=1e-8)
    B[st2 != 0, d:] = 0.0
    COUNTERS['irls_gene_iters'] += n_it + int(it2.sum())
    COUNTERS['fits'] += 1
    disp_out = nu_g if family == 'nb' else disp_glm
    fit = GlmFit(dev, B, dv, it2, st2, np.zeros(p, dtype=bool))
    last_fit_info.update(n_genes=int(p), n_refit=0, max_iter=int(it2.max(initial=0)), irls_gene_iters=n_it + int(it2.sum()),
                         route='fast-two-stage', n_stage1_cells=int(len(s1)))
    if _lazy:
        return B, fit, disp_out, offsets, None
    resid = fit.resid()
    if impute is not False:
        X_test = impute if isinstance(impute, np.ndarray) else X
        off_t = offsets
        if offset_test is not None and np.asarray(offset_test).shape[0] == X_test.shape[0]:
            off_t = np.asarray(offset_test, dtype=np.float64).ravel()
        elif offsets is not None and offsets.shape[0] != X_test.shape[0]:
            off_t = None
        eta0 = X_test @ B[:, :d].T
        if off_t is not None:
            eta0 = eta0 + off_t[:, None]
        Yhat_0 = np.repeat(np.exp(np.clip(eta0, -20, 20))[:, :, None], a, axis=2)
        Yhat_1 = np.exp(np.clip(eta0[:, :, None] + B[:, d:][None, :, :], -20, 20))
        Yhat = (Yhat_0, Yhat_1)
    else:
        Yhat = fit.fitted()
    return B, Yhat, disp_out, offsets, resid


def estimate_disp_auto(Y, X=None, A=None, Y_hat=None, disp_family='gaussian', offset=None, verbose=False, **kwargs):
    """Dispersion with a covariate-only Poisson design (``gcate_glm.py:467-474``).

    The reference returns ``None`` when crispyx is missing (and then runs with
    ``disp = 1``); here the in-tree moments estimator always runs."""
    X_disp = X if X is not None else np.ones((np.asarray(Y).shape[0], 1))
    return estimate_disp(Y, X_disp, offset=offset, disp_family='poisson', verbose=verbose, **kwargs)


def estimate_disp_fast(Y, X, *, A=None, offset=None, method="moments"):
    return estimate_disp(Y, X, A=A, offset=offset, disp_family='poisson')


def _compute_poisson_deviance_residuals(Y, mu):
    """Signed Poisson deviance residuals (host helper for small arrays)."""
    Y = np.asarray(Y, dtype=np.float64)
    mu = np.maximum(mu, 1e-10)
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.where(Y > 0, Y * np.log(np.maximum(Y, 1e-10) / mu) - (Y - mu), mu)
    r = np.sign(Y - mu) * np.sqrt(np.maximum(2.0 * term, 0.0))
    r[~np.isfinite(r)] = 0.0
    return r


def _compute_nb_deviance_residuals(Y, mu, alpha):
    """Signed NB deviance residuals with dispersion ``alpha = 1/size`` (host helper)."""
    Y = np.asarray(Y, dtype=np.float64)
    mu = np.maximum(mu, 1e-10)
    size = 1.0 / np.maximum(alpha, 1e-10)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = np.where(Y > 0, Y * np.log(np.maximum(Y, 1e-10) / mu), 0.0)
        t2 = (Y + size[None, :]) * np.log((Y + size[None, :]) / (mu + size[None, :]))
    r = np.sign(Y - mu) * np.sqrt(np.maximum(2.0 * (t1 - t2), 0.0))
    r[~np.isfinite(r)] = 0.0
    return r
