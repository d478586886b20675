"""CPU restatement of the per-gene GLM the reference delegates to statsmodels.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: statsmodels
is not vendored in ``/root/reference`` and is not installed here
(``environment.yaml:22`` says ``statsmodels>=0.14``); this file restates the
published algorithm of statsmodels 0.14 ``GLM._fit_irls`` /
``families.{Poisson,NegativeBinomial}`` / ``_MinimalWLS.fit(method='lstsq')``
and is anchored on the reference call sites:

* ``causarray/gcate_glm.py:187-191``  family construction (NB alpha = 1/disp)
* ``causarray/gcate_glm.py:193-236``  ``fit_model``: Poisson switch when
  ``disp > thres_disp``, ``sm.GLM(...).fit(maxiter=...)``, the |beta| guards
  (50 on the first ``d`` columns, 10 on the rest), ``resid_deviance``,
  ``predict`` on ``[X, 0]`` and ``[X, e_k]``.
* ``causarray/gcate_glm.py:275-308``  method-of-moments dispersion.

Algorithm restated (log link, scale 1, no freq/var weights):

    mu0  = (y + mean(y)) / 2 ; eta = log(mu0)            # family.starting_mu
    dev0 = deviance(y, mu0)
    repeat up to maxiter:
        w = mu            (Poisson)   |  mu / (1 + alpha mu)   (NB)
        z = eta + (y - mu) / mu - offset
        beta = lstsq(sqrt(w) X, sqrt(w) z)
        eta = X beta + offset ; mu = exp(eta)
        dev = deviance(y, mu)
        stop when |dev - dev_prev| <= 1e-8                 # atol=tol, rtol=0

Poisson unit deviance 2[y log(clip(y/mu, eps)) - (y - mu)];
NB unit deviance 2[y log(clip(y/mu, eps)) - (y + 1/alpha) log((y + 1/alpha)/(mu + 1/alpha))].
Deviance residual sign(y - mu) sqrt(max(d_i, 0)).
"""
import os

import numpy as np

FLOAT_EPS = np.finfo(float).eps


class _Family:
    name = None
    alpha = None


class _Poisson(_Family):
    name = "poisson"


class _NegativeBinomial(_Family):
    name = "nb"

    def __init__(self, alpha=1.0):
        self.alpha = float(alpha)


class _Gaussian(_Family):
    name = "gaussian"


class families:  # mirrors ``sm.families`` attribute access
    Poisson = _Poisson
    NegativeBinomial = _NegativeBinomial
    Gaussian = _Gaussian


def unit_deviance(y, mu, alpha=None):
    """Per-observation deviance contributions (statsmodels ``_resid_dev``)."""
    y = np.asarray(y, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ymu = np.clip(y / mu, FLOAT_EPS, np.inf)
        d = y * np.log(ymu)
        if alpha is None:
            d = d - (y - mu)
        else:
            ya = y + 1.0 / alpha
            ma = mu + 1.0 / alpha
            d = d - ya * np.log(ya / ma)
    return 2.0 * d


def resid_deviance(y, mu, alpha=None):
    d = unit_deviance(y, mu, alpha)
    return np.sign(y - mu) * np.sqrt(np.clip(d, 0.0, np.inf))


def irls(y, X, offset=None, alpha=None, maxiter=100, tol=1e-8):
    """One gene.  ``alpha`` None -> Poisson, else NB with that alpha.

    Returns ``(beta, mu, dev, n_iter, converged)``.  Raises ``ValueError`` on
    non-finite weights / working response, like ``_MinimalWLS``.
    """
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n = y.shape[0]
    off = np.zeros(n) if offset is None else np.asarray(offset, dtype=np.float64)
    mu = (y + y.mean()) / 2.0
    eta = np.log(mu)
    dev_prev = np.sum(unit_deviance(y, mu, alpha))
    if np.isnan(dev_prev):
        raise ValueError("first deviance is nan")
    beta = np.zeros(X.shape[1])
    converged = False
    it = 0
    for it in range(1, maxiter + 1):
        w = mu if alpha is None else mu / (1.0 + alpha * mu)
        z = eta + (y - mu) / mu - off
        sw = np.sqrt(w)
        if not np.all(np.isfinite(sw)):
            raise ValueError("non-finite weights")
        wz = sw * z
        if not np.all(np.isfinite(wz)):
            raise ValueError("non-finite working response")
        beta = np.linalg.lstsq(sw[:, None] * X, wz, rcond=-1)[0]
        eta = X @ beta + off
        with np.errstate(over="ignore"):
            mu = np.exp(eta)
        dev = np.sum(unit_deviance(y, mu, alpha))
        if abs(dev - dev_prev) <= tol:
            converged = True
            dev_prev = dev
            break
        dev_prev = dev
    return beta, mu, dev_prev, it, converged


def _lasso_quadratic(G, r, pen, b_start):
    """Minimiser of ``1/2 b^T G b - r^T b + sum_k pen_k |b_k|`` (G symmetric positive definite):
    active-set rounds on the sign pattern (exact linear solves), cyclic coordinate descent if the
    pattern does not settle."""
    k = r.size
    free = pen == 0
    b = np.linalg.solve(G, r) if b_start is None else b_start.copy()
    sg = np.sign(b)
    for _ in range(10):
        act = free | (sg != 0)
        rhs = r - np.where(free, 0.0, pen * sg)
        b = np.zeros(k)
        if act.any():
            b[act] = np.linalg.solve(G[np.ix_(act, act)], rhs[act])
        changed = False
        wrong = act & ~free & (np.sign(b) != sg)
        if wrong.any():
            sg = np.where(wrong, 0.0, sg)
            changed = True
        res = r - G @ b
        wake = ~act & (np.abs(res) > pen * (1.0 + 1e-12))
        if wake.any():
            sg = np.where(wake, np.sign(res), sg)
            changed = True
        if not changed:
            return b
    for sweep in range(20000):
        chg = 0.0
        for j in range(k):
            rj = r[j] - G[j] @ b + G[j, j] * b[j]
            nw = np.sign(rj) * max(abs(rj) - pen[j], 0.0) / G[j, j]
            chg = max(chg, abs(nw - b[j]))
            b[j] = nw
        if chg <= 1e-14 * (1.0 + np.abs(b).max()):
            break
    return b


def lasso_glm(y, X, offset, alpha_nb, pen, tol=1e-12, maxiter=500, step_tol=1e-11):
    """Minimiser of ``-loglike(b) / n + sum_k pen_k |b_k|`` for the log-link Poisson (``alpha_nb`` None)
    or NB2 (``var = mu + alpha_nb mu^2``) GLM, by proximal IRLS: each step minimises the weighted
    least-squares model plus the L1 term exactly (``_lasso_quadratic``)."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n, k = X.shape
    pen = np.broadcast_to(np.asarray(pen, dtype=np.float64), (k,)) * n
    off = 0.0 if offset is None else np.asarray(offset, dtype=np.float64)
    mu = (y + y.mean()) / 2.0
    eta = np.log(mu)
    beta = None
    obj_prev = np.inf
    for it in range(maxiter):
        w = mu if alpha_nb is None else mu / (1.0 + alpha_nb * mu)
        z = eta + (y - mu) / mu - off
        G = X.T @ (w[:, None] * X)
        r = X.T @ (w * z)
        beta_prev = beta
        beta = _lasso_quadratic(G, r, pen, beta)
        eta = X @ beta + off
        mu = np.exp(eta)
        obj = 0.5 * np.sum(unit_deviance(y, mu, alpha_nb)) + np.sum(pen * np.abs(beta))
        # the objective is flat along weakly identified columns (a relative change of 1e-10 still
        # leaves their coefficients ~1e-4 from the minimiser): also wait for the update to vanish
        step = np.inf if beta_prev is None else np.max(np.abs(beta - beta_prev) / (1.0 + np.abs(beta)))
        if abs(obj - obj_prev) <= tol * max(1.0, abs(obj)) and step <= step_tol:
            break
        obj_prev = obj
    return beta, mu


def lasso_kkt_violation(y, X, offset, alpha_nb, pen, beta):
    """Largest violation of the optimality conditions of ``lasso_glm``'s objective at ``beta``:
    ``g_k + pen_k sign(b_k) = 0`` where ``b_k != 0`` and ``|g_k| <= pen_k`` where ``b_k = 0``,
    ``g = -X^T [(y - mu) / (1 + alpha mu)] / n``."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n, k = X.shape
    pen = np.broadcast_to(np.asarray(pen, dtype=np.float64), (k,))
    eta = X @ beta + (0.0 if offset is None else offset)
    mu = np.exp(eta)
    q = 1.0 if alpha_nb is None else 1.0 / (1.0 + alpha_nb * mu)
    g = -(X.T @ ((y - mu) * q)) / n
    viol = np.where(beta != 0, np.abs(g + pen * np.sign(beta)), np.maximum(np.abs(g) - pen, 0.0))
    return float(viol.max())


class _Results:
    def __init__(self, model, beta, mu, n_iter, converged):
        self.model = model
        self.params = beta
        self.mu = mu
        self.n_iter = n_iter
        self.converged = converged

    @property
    def resid_deviance(self):
        return resid_deviance(self.model.endog, self.mu, self.model.family.alpha)

    def predict(self, exog, offset=None):
        eta = np.asarray(exog) @ self.params
        if offset is not None:
            eta = eta + offset
        return np.exp(eta)


def _count_iterations(it):
    """Work accounting for bench.py's CPU arm: with ``CA_ORACLE_ITER_DIR`` set, every process
    appends the IRLS iteration count of each fit to its own file there (the reference fans the
    per-gene fits out over joblib worker processes, so a module-level counter would not do)."""
    d = os.environ.get("CA_ORACLE_ITER_DIR")
    if d:
        with open(os.path.join(d, "iters.%d" % os.getpid()), "a") as f:
            f.write("%d\n" % it)


class SMGLM:
    """Minimal ``sm.GLM`` stand-in: ``SMGLM(y, X, family=..., offset=...).fit(maxiter=)``."""

    def __init__(self, endog, exog, family=None, offset=None):
        self.endog = np.asarray(endog, dtype=np.float64)
        self.exog = np.asarray(exog, dtype=np.float64)
        self.family = family if family is not None else _Gaussian()
        self.offset = offset

    def fit(self, maxiter=100, **kw):
        if self.family.name == "gaussian":
            beta = np.linalg.lstsq(self.exog, self.endog - (0 if self.offset is None else self.offset), rcond=-1)[0]
            mu = self.exog @ beta + (0 if self.offset is None else self.offset)
            res = _Results(self, beta, mu, 1, True)
            res.predict = lambda exog, offset=None: np.asarray(exog) @ beta + (0 if offset is None else offset)
            return res
        beta, mu, dev, it, conv = irls(self.endog, self.exog, self.offset, self.family.alpha, maxiter=maxiter)
        _count_iterations(it)
        return _Results(self, beta, mu, it, conv)

    def fit_regularized(self, alpha=0.0, cnvrg_tol=1e-5, L1_wt=1.0, **kw):
        """statsmodels ``GLM.fit_regularized`` (elastic net, ``statsmodels/base/elastic_net.py``) as
        called at ``causarray/gcate_glm.py:209``: minimise ``-loglike / nobs + sum_k alpha_k |b_k|``
        (``L1_wt = 1``).  statsmodels runs cyclic coordinate descent on the likelihood from zeros and
        stops when no coefficient moves by more than ``cnvrg_tol``; the objective is convex, so this
        restatement returns its minimiser (``lasso_glm``) -- equal to statsmodels' answer up to that
        tolerance.  PARITY UNPINNED (statsmodels is absent here); the optimality conditions are
        checked instead (``tests/test_oracle_golden.py``)."""
        if L1_wt != 1.0 or self.family.name == "gaussian":
            raise NotImplementedError("only the lasso fallback of the Poisson / NB fits is restated")
        beta, mu = lasso_glm(self.endog, self.exog, self.offset, self.family.alpha, alpha)
        return _Results(self, beta, mu, 0, True)


def fit_glm_original(Y, X, A=None, family="poisson", disp_glm=None, offset=None,
                     maxiter=1000, thres_disp=100.0, impute=False):
    """Restatement of ``fit_glm`` (``causarray/gcate_glm.py:95-272``), statsmodels path.

    Returns ``B (p, d[+a]), Yhat, disp_glm, offset, resid (n, p), status (p,)``
    where ``Yhat`` is ``(n, p)`` or ``(Yhat0, Yhat1)`` each ``(n, p, a)`` when
    ``impute`` is not False, and ``status[j]`` is 0 for a clean IRLS fit and 1
    when the reference takes its regularised fallback (guards at ``:205``) -- those
    genes get the lasso minimiser (``lasso_glm``) and zero residuals.
    """
    Y = np.asarray(Y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n, p = Y.shape
    d = X.shape[1]
    if A is None:
        a = 1
        Xf = X
    else:
        A = np.asarray(A, dtype=np.float64)
        if A.ndim == 1:
            A = A[:, None]
        a = A.shape[1]
        Xf = np.c_[X, A]
    B = np.zeros((p, Xf.shape[1]))
    resid = np.zeros((n, p))
    status = np.zeros(p, dtype=np.int32)
    n_iter = np.zeros(p, dtype=np.int32)
    Yhat0 = np.zeros((n, p, a))
    Yhat1 = np.zeros((n, p, a))
    off = None if offset is None else np.asarray(offset, dtype=np.float64)
    for j in range(p):
        alpha = None
        if family == "nb" and not (disp_glm[j] > thres_disp):
            alpha = 1.0 / disp_glm[j]
        try:
            beta, mu, dev, it, conv = irls(Y[:, j], Xf, off, alpha, maxiter=maxiter)
            n_iter[j] = it
            if (not np.all(np.isfinite(beta)) or np.any(np.abs(beta[:d]) > 50) or np.any(np.abs(beta[d:]) > 10)):
                raise ValueError("GLM did not converge")
        except ValueError:
            # the reference's fit_regularized fallback (gcate_glm.py:209): lasso with alpha = 1e-4 on
            # the non-constant columns, zero deviance residuals (:210)
            status[j] = 1
            try:
                pen = np.where(np.all(Xf == Xf[0, :], axis=0), 0.0, 1e-4)
                beta, mu = lasso_glm(Y[:, j], Xf, off, alpha, pen, tol=1e-10)
                if not np.all(np.isfinite(beta)):
                    raise ValueError("lasso fallback failed")
            except (ValueError, np.linalg.LinAlgError):
                continue
            B[j] = beta
            if impute is not False and A is not None:
                eta0 = X @ beta[:d] + (0 if off is None else off)
                for k in range(a):
                    Yhat0[:, j, k] = np.exp(eta0)
                    Yhat1[:, j, k] = np.exp(eta0 + beta[d + k])
            else:
                Yhat0[:, j, :] = mu[:, None]
                Yhat1[:, j, :] = mu[:, None]
            continue
        B[j] = beta
        resid[:, j] = resid_deviance(Y[:, j], mu, alpha)
        if impute is not False and A is not None:
            eta0 = X @ beta[:d] + (0 if off is None else off)
            for k in range(a):
                Yhat0[:, j, k] = np.exp(eta0)
                Yhat1[:, j, k] = np.exp(eta0 + beta[d + k])
        else:
            Yhat0[:, j, :] = mu[:, None]
            Yhat1[:, j, :] = mu[:, None]
    Yhat = (Yhat0, Yhat1) if (impute is not False and A is not None) else Yhat0[:, :, 0]
    return B, Yhat, disp_glm, offset, resid, status, n_iter


def estimate_disp_moments(Y, Y_hat_pois, offset=None):
    """Method-of-moments NB size (``causarray/gcate_glm.py:275-308``).

    ``Y_hat_pois`` are the fitted means of the Poisson GLM on the count scale;
    both are divided by the size factor ``exp(offset)`` first.
    """
    Y = np.asarray(Y, dtype=np.float64)
    sf = 1.0 if offset is None else np.exp(np.asarray(offset, dtype=np.float64))[:, None]
    Yn = Y / sf
    Yh = np.clip(Y_hat_pois / sf, 0.0, np.max(Yn, axis=0))
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = np.mean((Yn - Yh) ** 2 - Yh, axis=0) / np.mean(Yh ** 2, axis=0)
        disp = 1.0 / np.clip(phi, 0.01, 100.0)
    disp[np.isnan(disp)] = 1.0
    return disp
