"""CPU restatement of one perturbation batch of ``gcate_lfc_batch`` end to end
(TEST INFRASTRUCTURE: checker for tests/, and the timed CPU baseline / ``--impl reference``
arm of bench.py).

Flow (reference lines): ``fit_gcate`` (``causarray/gcate.py:54-193``) = size factors
(``utils.py:179-219``) -> ``alter_min`` stage 1 with GLM + SVD initialisation
(``gcate_opt.py:293-334``) -> ``qr(Gamma)`` -> stage 2; then ``LFC``
(``DR_learner.py:272-493``) = logistic propensities (``DR_estimation.py:66-200``),
outcome GLM with MoM dispersion (``gcate_glm.py:95-308``), ``AIPW_mean``
(``DR_estimation.py:624-659``), the ``estimand`` closure (``DR_learner.py:403-482``) and
``bh_correction`` (``DR_inference.py:153-201``).

The alternating iteration runs in ``oracle/gcate_oracle.c`` (OpenMP over rows, as numba's
prange), the per-gene IRLS in ``oracle/glm.py`` fanned out over processes with joblib (as
``gcate_glm.py:239``).
"""
import os
import time

import numpy as np
import scipy.linalg
from scipy.sparse.linalg import svds

from . import altmin as om
from . import glm as og
from . import inference as oi


def comp_size_factor(counts):
    """``causarray/utils.py:197-213`` (method='geomeans')."""
    counts = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        logc = np.where(counts > 0, np.log(counts), np.nan)
        lg = np.nanmean(logc, axis=0)
    lg[np.all(counts == 0, axis=0)] = -np.inf
    mask = np.isfinite(lg)[None, :] & (counts > 0)
    diff = np.where(mask, logc - lg[None, :], np.nan)
    ls = np.nanmedian(diff, axis=1)
    return np.exp(ls - np.mean(ls))


def _svds_det(M, k):
    v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
    u, s, vt = svds(M, k=k, v0=v0, tol=1e-14)
    order = np.argsort(-s)
    u, s, vt = u[:, order], s[order], vt[order]
    for i in range(k):
        j = np.argmax(np.abs(u[:, i]))
        if u[j, i] < 0:
            u[:, i] *= -1
            vt[i] *= -1
    return u, s, vt


def _glm_chunk(Ychunk, X, A, family, disp, offset, maxiter, impute):
    return og.fit_glm_original(Ychunk, X, A, family=family, disp_glm=disp, offset=offset, maxiter=maxiter,
                               impute=impute)


def fit_glm_parallel(Y, X, A=None, family="poisson", disp_glm=None, offset=None, maxiter=1000, impute=False,
                     n_jobs=1):
    """Per-gene IRLS over ``n_jobs`` processes.  Returns ``B, resid, status, n_iter, mu``."""
    p = Y.shape[1]
    if n_jobs is None or n_jobs <= 1:
        parts = [_glm_chunk(Y, X, A, family, disp_glm, offset, maxiter, impute)]
        bounds = [(0, p)]
    else:
        from joblib import Parallel, delayed
        nchunk = min(p, n_jobs * 2)
        edges = np.linspace(0, p, nchunk + 1).astype(int)
        bounds = [(edges[i], edges[i + 1]) for i in range(nchunk) if edges[i + 1] > edges[i]]
        parts = Parallel(n_jobs=n_jobs)(
            delayed(_glm_chunk)(np.ascontiguousarray(Y[:, lo:hi]), X, A, family,
                                None if disp_glm is None else disp_glm[lo:hi], offset, maxiter, impute)
            for lo, hi in bounds)
    B = np.concatenate([q[0] for q in parts], axis=0)
    resid = np.concatenate([q[4] for q in parts], axis=1)
    status = np.concatenate([q[5] for q in parts])
    n_iter = np.concatenate([q[6] for q in parts])
    if impute is False or A is None:
        mu = np.concatenate([q[1] for q in parts], axis=1)
    else:
        mu = (np.concatenate([q[1][0] for q in parts], axis=1), np.concatenate([q[1][1] for q in parts], axis=1))
    return B, resid, status, n_iter, mu


def estimand_lfc(etas, A, usevar="unequal", thres_min=1e-2, thres_diff=1e-2, eps_var=1e-4):
    """``causarray/DR_learner.py:403-482``."""
    e0, e1 = etas[..., 0], etas[..., 1]
    m0 = np.mean(e0, axis=0, dtype=np.float64)
    m1 = np.mean(e1, axis=0, dtype=np.float64)
    est = np.isfinite(m0) & np.isfinite(m1) & (m0 > 0) & (m1 > 0)
    t0 = np.where(est, np.maximum(m0, thres_diff), 1.0)
    t1 = np.where(est, np.maximum(m1, thres_diff), 1.0)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        tau = np.log(t1 / t0)
        IF = e1 / t1[None, :] - e0 / t0[None, :]
    df = None
    if usevar == "pooled":
        var = (np.var(IF, axis=0, ddof=1) + eps_var) / IF.shape[0]
    else:
        n0, n1 = int(np.sum(A == 0)), int(np.sum(A == 1))
        with np.errstate(invalid="ignore", divide="ignore"):
            v0 = (np.var(IF[A == 0], axis=0, ddof=1) + eps_var) / n0
            v1 = (np.var(IF[A == 1], axis=0, ddof=1) + eps_var) / n1
            var = v0 + v1
            df = (v0 + v1) ** 2 / (v0 ** 2 / (n0 - 1) + v1 ** 2 / (n1 - 1))
    idx = ~est | (np.maximum(m0, m1) < thres_min) | (np.abs(m1 - m0) < thres_diff)
    tau[idx] = 0.0
    var[idx] = np.inf
    if df is not None:
        df[idx] = np.nan
    return tau, var, df, m0, m1, est


def fit_gcate_cpu(Y, X, A, r, family, disp, es1=None, es2=None, c1=0.05, n_jobs=1, use_c=True, stats=None):
    """``fit_gcate`` with explicit dispersion: returns dict with X_U, B_Gamma, size_factor, hist1, hist2."""
    from . import cport
    stats = {} if stats is None else stats
    Y = np.asarray(Y, dtype=np.float64)
    Xf = np.c_[X, A]
    n, p = Y.shape
    d, a = Xf.shape[1], A.shape[1]
    t0 = time.perf_counter()
    sf = comp_size_factor(Y)
    off = np.log(sf)
    stats["t_size_factor"] = time.perf_counter() - t0
    Q1 = scipy.linalg.qr(Xf, mode="economic")[0]
    t0 = time.perf_counter()
    B1, resid, st1, it1, _ = fit_glm_parallel(Y, Xf, None, family, disp, off, maxiter=100, n_jobs=n_jobs)
    u, s, vt = _svds_det(resid, r)
    A0 = np.c_[Xf, u - Q1 @ (Q1.T @ u)]
    B0, _, st2, it2, _ = fit_glm_parallel(Y, A0, None, family, disp, off, maxiter=100, n_jobs=n_jobs)
    E = A0[:, -r:] @ B0[:, -r:].T
    u, s, vt = _svds_det(E, r)
    A0[:, d:] = u * s[None, :] ** 0.5
    B0[:, d:] = vt.T * s[None, :] ** 0.5
    stats["t_init"] = time.perf_counter() - t0
    stats["irls_gene_iters"] = stats.get("irls_gene_iters", 0) + int(it1.sum()) + int(it2.sum())
    upd = None
    if use_c:
        Yn = Y / sf[:, None]
        nu = np.asarray(disp, dtype=np.float64).reshape(1, -1)
        prob = cport.Problem(Yn, om.log_h_terms(Yn, nu, family, 100.0))

        def upd(Yn_, A_, B_, *args, **kw):
            return cport.update_with_mask(Yn_, A_, B_, *args, _prob=prob, **kw)
    t0 = time.perf_counter()
    r1 = alter_min_loop_c(Y, Xf, r, A0, B0, family, disp, sf, 1, kwargs_es=es1, update_fn=upd)
    Q2 = scipy.linalg.qr(r1["B_Gamma"][:, -r:], mode="economic")[0]
    r2 = alter_min_loop_c(Y, Xf, r, r1["X_U"].copy(), r1["B_Gamma"].copy(), family, disp, sf, 2, P2=Q2, lam=c1, a=a,
                          kwargs_es=es2, update_fn=upd)
    stats["t_altmin"] = time.perf_counter() - t0
    stats["alt_iters"] = (len(r1["hist"]) - 1) + (len(r2["hist"]) - 1)
    return dict(X_U=r2["X_U"], B_Gamma=r2["B_Gamma"], U=r2["X_U"][:, d:], size_factor=sf, hist1=r1["hist"],
                hist2=r2["hist"], s1_A=r1["X_U"], s1_B=r1["B_Gamma"])


def alter_min_loop_c(Y, X, r, A0, B0, family, disp, sf, stage, P2=None, lam=0.0, a=None, kwargs_es=None,
                     update_fn=None):
    """``oracle.altmin.alter_min_loop`` with a pluggable iteration kernel (C port by default)."""
    if update_fn is None:
        return om.alter_min_loop(Y, X, r, A0, B0, family, disp, sf, stage, P2=P2, lam=lam, a=a, kwargs_es=kwargs_es)
    orig = om.update_with_mask
    om.update_with_mask = update_fn
    try:
        return om.alter_min_loop(Y, X, r, A0, B0, family, disp, sf, stage, P2=P2, lam=lam, a=a, kwargs_es=kwargs_es)
    finally:
        om.update_with_mask = orig


def lfc_cpu(Y, W, A, W_A, family, offset, usevar="unequal", disp=None, n_jobs=1, stats=None):
    """``LFC`` (K = 1, logistic propensities, statsmodels-path outcome model)."""
    from sklearn.linear_model import LogisticRegression
    stats = {} if stats is None else stats
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    a = A.shape[1]
    sf = np.exp(offset)
    d = W.shape[1]
    t0 = time.perf_counter()
    ctrl = A.sum(axis=1) == 0
    pi_raw = np.zeros((n, a))
    for j in range(a):
        elig = ctrl | (A[:, j] == 1)
        clf = LogisticRegression(fit_intercept=False, C=1.0, class_weight="balanced", random_state=0)
        pi_raw[:, j] = clf.fit(W_A[elig], A[elig, j]).predict_proba(W_A)[:, 1]
    pi = np.clip(pi_raw, 0.01, 0.99)
    stats["t_propensity"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Xf = np.c_[W, A]
    it_sum = 0
    if family == "nb" and disp is None:
        _, _, _, itp, mu_p = fit_glm_parallel(Y, Xf, None, "poisson", None, offset, maxiter=1000, n_jobs=n_jobs)
        disp = og.estimate_disp_moments(Y, mu_p, offset)
        it_sum += int(itp.sum())
    # (treatments passed separately: their coefficients carry the tighter guard |beta| > 10 of
    # gcate_glm.py:205, as in the reference's cross_fitting -> fit_glm(Y, X, A) call)
    B, _, status, itn, _ = fit_glm_parallel(Y, W, A, family, disp, offset, maxiter=1000, n_jobs=n_jobs)
    it_sum += int(itn.sum())
    stats["t_glm_lfc"] = time.perf_counter() - t0
    stats["irls_gene_iters"] = stats.get("irls_gene_iters", 0) + it_sum
    t0 = time.perf_counter()
    # status == 1 genes: guard tripped -- lasso fallback coefficients (oracle/glm.py::lasso_glm)
    eta0 = W @ B[:, :d].T + offset[:, None]
    with np.errstate(over="ignore"):
        Yh0 = np.minimum(np.exp(eta0), 1e5)
    tau = np.zeros((a, p))
    var = np.zeros((a, p))
    dfe = np.zeros((a, p))
    m0 = np.zeros((a, p))
    m1 = np.zeros((a, p))
    est = np.zeros((a, p), dtype=bool)
    pv = np.zeros((a, p))
    qv = np.zeros((a, p))
    for j in range(a):
        cells = ctrl | (A[:, j] == 1)
        with np.errstate(over="ignore"):
            Yh1 = np.minimum(np.exp(eta0[cells] + B[:, d + j][None, :]), 1e5)
        Aj = A[cells, j]
        Yc = Y[cells]
        w0 = ((1 - Aj) / (1 - pi[cells, j]))[:, None]
        w1 = (Aj / pi[cells, j])[:, None]
        etas = np.stack([w0 * (Yc - Yh0[cells]) + Yh0[cells], w1 * (Yc - Yh1) + Yh1], axis=-1)
        etas /= sf[cells][:, None, None]
        tau[j], var[j], df_j, m0[j], m1[j], est[j] = estimand_lfc(etas, Aj, usevar=usevar)
        with np.errstate(invalid="ignore", divide="ignore"):
            std = np.sqrt(var[j])
            t = tau[j] / std
        t[np.isinf(std)] = np.nan
        pv[j], qv[j], _, _ = oi.bh_correction(t, df=df_j)
        dfe[j] = np.nan if df_j is None else df_j
    stats["t_aipw"] = time.perf_counter() - t0
    return dict(tau=tau, var=var, df=dfe, mean0=m0, mean1=m1, estimable=est, pvalue=pv, padj=qv, pi=pi, B=B,
                status=status, disp=disp)


def gcate_lfc_one_batch_cpu(Y, X, A, X_A, r, family="nb", disp=None, es1=None, es2=None, usevar="unequal",
                            n_jobs=None, stats=None):
    """One perturbation batch of ``gcate_lfc_batch`` on the CPU; returns ``(gcate, lfc, stats)``."""
    stats = {} if stats is None else stats
    if n_jobs is None:
        n_jobs = os.cpu_count() or 1
    t0 = time.perf_counter()
    g = fit_gcate_cpu(Y, X, A, r, family, disp, es1=es1, es2=es2, n_jobs=n_jobs, stats=stats)
    W = np.c_[X, g["U"]]
    WA = np.c_[X_A, g["U"]]
    l = lfc_cpu(Y, W, A, WA, family, np.log(g["size_factor"]), usevar=usevar, n_jobs=n_jobs, stats=stats)
    stats["t_total"] = time.perf_counter() - t0
    n, p = Y.shape
    stats["cell_gene_updates"] = n * p * stats["alt_iters"] + n * stats["irls_gene_iters"]
    return g, l, stats
