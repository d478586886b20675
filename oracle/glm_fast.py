"""CPU restatement of the two-stage "fast" GLM route (TEST INFRASTRUCTURE).

Reference: ``causarray/nb_glm_fast.py:373-594`` (``_fit_glm_fast_per_perturbation``).  The arithmetic
inside its stages is crispyx's ``NBGLMBatchFitter`` -- an un-vendored, unpinned third-party package
(``setup.cfg:30``) that is absent from this image: PARITY UNPINNED.  What is restated here is the
structure the reference spells out in-tree, with the statsmodels-form IRLS of ``oracle/glm.py`` inside:

* stage 1 (``:395-462``): covariate-only design on at most 3 000 cells
  (``default_rng(random_state).choice(n, 3000, replace=False)``, sorted); Poisson IRLS capped at
  5 (NB) / 10 (Poisson) iterations; method-of-moments dispersion from its fitted means,
  ``alpha = clip(., 1e-8, 100)``; for NB an NB IRLS capped at 5 iterations;
* stage 2 (``:500-567``): for every perturbation k a 1-D fit of the treatment effect on control +
  perturbed cells with the covariate offset ``X B_cov^T`` and the dispersion fixed, at most 50
  iterations.  Control cells have a zero design entry and drop out of the 1-D normal equation; the
  stopping rule used here looks at the deviance over ALL cells (one joint model with decoupled
  one-hot columns), which is what the device path does.
"""
import numpy as np

from . import glm as og

N_STAGE1 = 3000


def fit_glm_fast_two_stage(Y, X, A, family="nb", offset=None, maxiter=1000, random_state=0):
    """Returns ``B (p, d + a)``, ``nu (p,)`` (None for Poisson), ``n_iter_stage2 (p,)``."""
    Y = np.asarray(Y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    n, p = Y.shape
    d, a = X.shape[1], A.shape[1]
    off = np.zeros(n) if offset is None else np.asarray(offset, dtype=np.float64)
    s1 = np.arange(n) if n <= N_STAGE1 else np.sort(np.random.default_rng(random_state).choice(n, N_STAGE1, replace=False))
    cap = 5 if family == "nb" else 10
    B_cov = np.zeros((p, d))
    nu = np.ones(p) if family == "nb" else None
    for j in range(p):
        y = Y[s1, j]
        beta, mu, _, _, _ = og.irls(y, X[s1], off[s1], None, maxiter=min(maxiter, cap))
        if family == "nb":
            disp = og.estimate_disp_moments(y[:, None], mu[:, None], off[s1])[0]
            al = float(np.clip(1.0 / disp, 1e-8, 100.0))
            if not np.isfinite(al):
                al = 1.0
            nu[j] = 1.0 / al
            beta, mu, _, _, _ = og.irls(y, X[s1], off[s1], al, maxiter=min(maxiter, 5))
        B_cov[j] = beta
    # stage 2: joint model with the covariate part as an offset; one-hot columns decouple
    B_trt = np.zeros((p, a))
    n_iter = np.zeros(p, dtype=int)
    grp = [np.where(A[:, k] == 1)[0] for k in range(a)]
    for j in range(p):
        y = Y[:, j]
        al = None if family != "nb" else 1.0 / nu[j]
        base = X @ B_cov[j] + off
        bt = np.zeros(a)
        dev_prev = None
        for it in range(1, min(maxiter, 50) + 2):
            eta = base + A @ bt
            mu = np.exp(eta)
            dev = float(np.sum(og.unit_deviance(y, mu, al)))
            if dev_prev is not None and abs(dev - dev_prev) <= 1e-8:
                n_iter[j] = it - 1
                break
            dev_prev = dev
            if it > min(maxiter, 50):
                n_iter[j] = min(maxiter, 50)
                break
            w = mu if al is None else mu / (1.0 + al * mu)
            z = (A @ bt) + (y - mu) / mu
            for k in range(a):
                g = grp[k]
                bt[k] = np.sum(w[g] * z[g]) / np.sum(w[g])
        B_trt[j] = bt
    return np.c_[B_cov, B_trt], nu, n_iter
