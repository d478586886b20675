"""CPU: the oracle (numpy restatement and the C/OpenMP port) against the golden vectors
produced by the REAL reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import altmin as om
from oracle import glm as og
from oracle import inference as oi


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_nll_grad(golden_dir, fam):
    g = _load(golden_dir, "kernels_%s.npz" % fam)
    nu = g["nu"].reshape(1, -1)
    Yn = g["Y"] / g["sf"][:, None]
    Ys = om.log_h_terms(Yn, nu, fam, float(g["thres_disp"]))
    np.testing.assert_allclose(Ys, g["Ys"], rtol=0, atol=1e-12)
    assert abs(om.nll(Yn, g["A0"], g["B0"], fam, nu, Ys, 100.0) - g["nll"]) < 1e-10
    np.testing.assert_allclose(om.grad(Yn, g["A0"], g["B0"], fam, nu, 100.0), g["grad_B"], atol=1e-13)
    np.testing.assert_allclose(om.grad(Yn.T, g["B0"], g["A0"], fam, nu.T, 100.0), g["grad_A"], atol=1e-13)


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("fam", ["nb", "poisson", "nb_clip"])
def test_update_with_mask(golden_dir, fam, impl):
    """(nb_clip: radius 2C = 0.02, so project_norm_ball -- causarray/gcate_opt.py:10-26 -- rescales 52 of
    60 cell gradients and 53 of 70 gene gradients of the golden problem)"""
    g = _load(golden_dir, "kernels_%s.npz" % fam)
    Cg = float(g["C"]) if "C" in g else None
    fam = fam.split("_")[0]
    if impl == "c":
        from oracle import cport
        upd = cport.update_with_mask
    else:
        upd = om.update_with_mask
    nu = g["nu"].reshape(1, -1)
    d, a = int(g["d"]), int(g["a"])
    Yn = g["Y"] / g["sf"][:, None]
    Ys = g["Ys"]
    C = Cg if Cg is not None else (1e3 if fam == "nb" else 1e5)
    A, B = g["A0"].copy(), g["B0"].copy()
    lam1 = np.zeros((g["Y"].shape[1], d))
    for it in range(3):
        f, A, B = upd(Yn, A, B, d, lam1, g["Q1"], None, fam, nu, Ys, 100.0, d, C, 0.1, 0.5, 20, 1e-4,
                      g["gene_active"], g["cell_active"], g["alpha_gene"])
        assert abs(f - g["s1_f"][it]) < 1e-9
        np.testing.assert_allclose(A, g["s1_A_%d" % it], atol=1e-12)
        np.testing.assert_allclose(B, g["s1_B_%d" % it], atol=1e-12)
    A2, B2 = g["s2_A_in"].copy(), g["s2_B_in"].copy()
    for it in range(3):
        f, A2, B2 = upd(Yn, A2, B2, d, g["lam2"], None, g["Q2"], fam, nu, Ys, 100.0, a, C, 0.1, 0.5, 20, 1e-4,
                        g["gene_active"], g["cell_active"], g["alpha_gene"])
        assert abs(f - g["s2_f"][it]) < 1e-9
        np.testing.assert_allclose(A2, g["s2_A_%d" % it], atol=1e-12)
        np.testing.assert_allclose(B2, g["s2_B_%d" % it], atol=1e-12)


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_alter_min_loop(golden_dir, fam):
    from oracle import cport, pipeline
    g = _load(golden_dir, "alter_min_%s.npz" % fam)
    Y, X = g["Y"], g["X"]
    r, a = int(g["r"]), int(g["a"])
    r1 = pipeline.alter_min_loop_c(Y, X, r, g["A0"], g["B0"], fam, g["disp"], g["sf"], 1,
                                   kwargs_es=dict(max_iters=30), update_fn=cport.update_with_mask)
    assert r1["n_iter"] == int(g["s1_n_iter"])
    np.testing.assert_allclose(r1["hist"], g["s1_hist"], rtol=1e-10)
    np.testing.assert_allclose(r1["X_U"], g["s1_A"], atol=1e-9)
    np.testing.assert_allclose(r1["B_Gamma"], g["s1_B"], atol=1e-9)
    assert r1["total_gene_updates"] == int(g["s1_gene_updates"])
    r2 = pipeline.alter_min_loop_c(Y, X, r, g["s1_A"], g["s1_B"], fam, g["disp"], g["sf"], 2, P2=g["Q2"], lam=0.05,
                                   a=a, kwargs_es=dict(max_iters=30), update_fn=cport.update_with_mask)
    assert r2["n_iter"] == int(g["s2_n_iter"])
    np.testing.assert_allclose(r2["hist"], g["s2_hist"], rtol=1e-10)
    np.testing.assert_allclose(r2["B_Gamma"], g["s2_B"], atol=1e-9)


@pytest.mark.parametrize("usevar", ["unequal", "pooled"])
def test_lfc_estimand(golden_dir, usevar):
    from oracle import pipeline
    g = _load(golden_dir, "lfc_%s.npz" % usevar)
    Y, W, A, pi, off = g["Y"], g["W"], g["A"], g["pi"], g["offset"]
    n, p = Y.shape
    a = A.shape[1]
    sf = np.exp(off)
    eta0 = W @ g["Bcov"].T + off[:, None]
    ctrl = A.sum(1) == 0
    tau = np.zeros((a, p))
    std = np.zeros((a, p))
    pv = np.zeros((a, p))
    qv = np.zeros((a, p))
    for j in range(a):
        cells = ctrl | (A[:, j] == 1)
        Yh0 = np.minimum(np.exp(eta0[cells]), 1e5)
        Yh1 = np.minimum(np.exp(eta0[cells] + g["Btrt"][:, j][None, :]), 1e5)
        Aj = A[cells, j]
        w0 = ((1 - Aj) / (1 - pi[cells, j]))[:, None]
        w1 = (Aj / pi[cells, j])[:, None]
        etas = np.stack([w0 * (Y[cells] - Yh0) + Yh0, w1 * (Y[cells] - Yh1) + Yh1], -1) / sf[cells][:, None, None]
        t, v, df, m0, m1, est = pipeline.estimand_lfc(etas, Aj, usevar=usevar)
        tau[j], std[j] = t, np.sqrt(v)
        with np.errstate(invalid="ignore", divide="ignore"):
            tv = t / np.sqrt(v)
        tv[np.isinf(std[j])] = np.nan
        pv[j], qv[j], _, _ = oi.bh_correction(tv, df=df)
    np.testing.assert_allclose(tau.ravel(), g["df_tau"], atol=1e-13)
    fin = np.isfinite(g["df_std"])
    np.testing.assert_allclose(std.ravel()[fin], g["df_std"][fin], rtol=1e-12)
    np.testing.assert_allclose(pv.ravel()[fin], g["df_pvalue"][fin], rtol=1e-10)
    np.testing.assert_allclose(qv.ravel()[fin], g["df_padj"][fin], rtol=1e-10)


def test_misc_golden(golden_dir):
    from oracle import pipeline
    g = _load(golden_dir, "misc.npz")
    np.testing.assert_allclose(pipeline.comp_size_factor(g["Y"]), g["size_factor"], rtol=1e-13)
    for pre, df in (("bh", g["bh_df"]), ("bhz", None)):
        out = oi.bh_correction(g["bh_t"].copy(), df=df)
        for got, name in zip(out, ["p", "q", "pa", "qa"]):
            np.testing.assert_allclose(got, g["%s_%s" % (pre, name)], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(og.resid_deviance(g["Y"], g["mu"], None), g["resid_pois"], atol=1e-10)
    np.testing.assert_allclose(og.resid_deviance(g["Y"], g["mu"], g["alpha"][None, :]), g["resid_nb"], atol=1e-9)
    np.testing.assert_allclose(og.estimate_disp_moments(g["Y"], g["mu"], np.log(g["size_factor"])), g["disp_mom"],
                               rtol=1e-12)


def test_irls_is_the_mle():
    """The IRLS restatement (parity unpinned vs statsmodels) reaches the MLE: cross-check
    against scipy's BFGS on the Poisson / NB log-likelihood and sklearn's PoissonRegressor."""
    import scipy.optimize
    from sklearn.linear_model import PoissonRegressor
    rng = np.random.default_rng(0)
    n = 400
    X = np.c_[np.ones(n), rng.standard_normal((n, 3))]
    beta = np.array([0.5, 0.3, -0.2, 0.1])
    off = rng.standard_normal(n) * 0.2
    mu = np.exp(X @ beta + off)
    y = rng.poisson(mu).astype(float)
    b, _, _, it, conv = og.irls(y, X, off, None, maxiter=100)
    assert conv
    sk = PoissonRegressor(alpha=0, fit_intercept=False, tol=1e-12, max_iter=1000).fit(
        X, y / np.exp(off), sample_weight=np.exp(off))
    np.testing.assert_allclose(b, sk.coef_, rtol=1e-5, atol=1e-6)
    alpha = 0.7
    ynb = rng.negative_binomial(1 / alpha, (1 / alpha) / (1 / alpha + mu)).astype(float)
    bn, _, _, _, conv = og.irls(ynb, X, off, alpha, maxiter=200)
    assert conv

    def negll(bb):
        m = np.exp(X @ bb + off)
        return -np.sum(ynb * np.log(alpha * m / (1 + alpha * m)) - np.log1p(alpha * m) / alpha)
    opt = scipy.optimize.minimize(negll, bn * 0, method="BFGS", options=dict(gtol=1e-9))
    np.testing.assert_allclose(bn, opt.x, rtol=1e-4, atol=1e-5)


def test_poisson_deviance_known_answers():
    """Known answers of the reference's own tests (tests/test_nb_glm_fast.py:39-58)."""
    mu = np.array([[2.0, 3.0]])
    np.testing.assert_allclose(og.resid_deviance(np.zeros((1, 2)), mu, None), -np.sqrt(2 * mu))
    np.testing.assert_allclose(og.resid_deviance(mu.copy(), mu, None), 0.0, atol=1e-12)


def test_oracle_lasso_fallback_is_the_minimiser():
    """The restated ``fit_regularized`` fallback (lasso objective of statsmodels' elastic net):
    optimality conditions hold at its answer and the unpenalised MLE violates them by the penalty."""
    from oracle import glm as og
    rng = np.random.default_rng(0)
    n, k = 300, 5
    X = np.c_[np.ones(n), rng.standard_normal((n, k - 2)) * 0.4, rng.standard_normal(n) * 0.02]
    beta = np.array([1.0, 0.3, -0.2, 0.1, 25.0])
    y = rng.poisson(np.exp(X @ beta)).astype(float)
    pen = np.r_[0.0, np.full(k - 1, 1e-4)]
    for a_nb in (None, 0.8):
        b, mu = og.lasso_glm(y, X, None, a_nb, pen)
        assert og.lasso_kkt_violation(y, X, None, a_nb, pen, b) < 1e-8
        b_mle = og.irls(y, X, None, a_nb, maxiter=200)[0]
        assert abs(og.lasso_kkt_violation(y, X, None, a_nb, pen, b_mle) - 1e-4) < 2e-6
        fam = og.families.Poisson() if a_nb is None else og.families.NegativeBinomial(alpha=a_nb)
        res = og.SMGLM(y, X, family=fam, offset=None).fit_regularized(alpha=pen, cnvrg_tol=1e-5)
        np.testing.assert_allclose(res.params, b)
