"""GPU parity of the S2 (per-gene GLM) and S3 (AIPW / LFC) seams."""
import os

import numpy as np
import pytest

from oracle import glm as og

pytestmark = pytest.mark.gpu


def _glm_problem(fam, seed, n_ctrl=120, p=90):
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=n_ctrl, n_perts=3, cells_per_pert=40, p=p, r=2, d_cov=2, family=fam, seed=seed,
                          base_lo=0.0, base_hi=2.5)
    rng = np.random.default_rng(seed)
    Y = sim["Y"]
    X = np.c_[sim["X"], sim["A"], sim["U"] * 0.3]
    off = rng.standard_normal(Y.shape[0]) * 0.2
    return Y, X, off, sim["disp"]


@pytest.mark.parametrize("fam", ["poisson", "nb"])
def test_glm_fit_vs_oracle(fam):
    from causarray_b200.device import GlmDevice
    Y, X, off, disp = _glm_problem(fam, 3)
    n, p = Y.shape
    disp = disp.copy()
    disp[::7] = 500.0   # Poisson switch inside the NB family
    B_ref, Yhat_ref, _, _, resid_ref, status_ref, it_ref = og.fit_glm_original(
        Y, X, None, family=fam, disp_glm=disp, offset=off, maxiter=100)
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    B, dv, n_iter, status = dev.fit(fam, X, off, disp if fam == "nb" else None, maxiter=100, tol=1e-8, d_guard=2)
    ok = (status == 0) & (status_ref == 0)
    print(fam, "ok genes", ok.sum(), "of", p, "max |dB|", np.abs(B[ok] - B_ref[ok]).max(), "iters", n_iter[:10],
          it_ref[:10], "iter mismatches", int((n_iter[ok] != it_ref[ok]).sum()))
    assert ok.sum() >= p - 2
    np.testing.assert_allclose(B[ok], B_ref[ok], rtol=1e-7, atol=1e-8)
    assert (n_iter[ok] != it_ref[ok]).sum() <= max(1, p // 50)
    mu = dev.fitted()
    np.testing.assert_allclose(mu[:, ok], Yhat_ref[:, ok], rtol=1e-7)
    resid = dev.resid()
    np.testing.assert_allclose(resid[:, ok], resid_ref[:, ok], rtol=1e-6, atol=1e-7)
    dm = dev.disp_moments(off)
    dm_ref = og.estimate_disp_moments(Y, mu, off)
    np.testing.assert_allclose(dm, dm_ref, rtol=1e-9)


def test_glm_reload_counts_same_handle():
    """Replacing the counts of a handle (same shape, same design) must not reuse the group-sorted copy the IRLS
    kernel made of the PREVIOUS counts (the device buffer, hence its address, is the same)."""
    from causarray_b200.device import GlmDevice
    Y, X, off, disp = _glm_problem("nb", 3)
    n, p = Y.shape
    h = p // 2
    Y1, Y2 = np.ascontiguousarray(Y[:, :h]), np.ascontiguousarray(Y[:, h:2 * h])
    dev = GlmDevice(n, h)
    dev.load_counts(Y1)
    B1 = dev.fit("nb", X, off, disp[:h], maxiter=100, tol=1e-8, d_guard=-1)[0]
    dev.load_counts(Y2)
    B2 = dev.fit("nb", X, off, disp[h:2 * h], maxiter=100, tol=1e-8, d_guard=-1)[0]
    fresh = GlmDevice(n, h)
    fresh.load_counts(Y2)
    B2f = fresh.fit("nb", X, off, disp[h:2 * h], maxiter=100, tol=1e-8, d_guard=-1)[0]
    assert not np.array_equal(B1, B2)
    np.testing.assert_array_equal(B2, B2f)


@pytest.mark.parametrize("kind", ["fractional", "large"])
def test_glm_log_y_table_fallback(kind):
    """The IRLS kernel reads log y of small integer counts from a table; responses that are not small
    non-negative integers (normalised / fractional values, counts of 256 and more) take the computed
    path -- both against the oracle."""
    from causarray_b200.device import GlmDevice
    Y, X, off, disp = _glm_problem("nb", 5)
    Y = np.asarray(Y, dtype=np.float64).copy()
    if kind == "fractional":
        Y[::3] += 0.5
    else:
        Y[:, ::4] *= 300.0      # every fourth gene far above the table's range
        assert Y.max() >= 256
    n, p = Y.shape
    B_ref, _, _, _, _, status_ref, it_ref = og.fit_glm_original(Y, X, None, family="nb", disp_glm=disp, offset=off,
                                                             maxiter=100)
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    B, dv, n_iter, status = dev.fit("nb", X, off, disp, maxiter=100, tol=1e-8, d_guard=2)
    ok = (status == 0) & (status_ref == 0)
    assert ok.sum() >= p - 4
    np.testing.assert_allclose(B[ok], B_ref[ok], rtol=1e-7, atol=1e-8)
    assert (n_iter[ok] != it_ref[ok]).sum() <= max(1, p // 50)


@pytest.mark.parametrize("fam,d_cov,r", [("nb", 2, 4), ("poisson", 3, 8)])
def test_glm_orthogonal_block_layout(fam, d_cov, r):
    """One-hot treatment block: the tile-skipping internal layout (dense part of one / two tiles)
    against the plain layout and the oracle."""
    from causarray_b200.device import GlmDevice
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=150, n_perts=12, cells_per_pert=25, p=70, r=r, d_cov=d_cov, family=fam, seed=11,
                          base_lo=0.0, base_hi=2.5)
    Y = sim["Y"]
    X = np.c_[sim["X"], sim["U"] * 0.3, sim["A"]]
    assert X.shape[1] >= 17 and set(np.unique(sim["A"])) <= {0.0, 1.0}
    n, p = Y.shape
    off = np.random.default_rng(5).standard_normal(n) * 0.2
    disp = sim["disp"] if fam == "nb" else None
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    out = {}
    for tag in ("ortho", "plain"):
        if tag == "plain":
            os.environ["CA_GLM_NO_ORTHO"] = "1"
        try:
            B, dv, n_iter, status = dev.fit(fam, X, off, disp, maxiter=100, tol=1e-8, d_guard=2)
            out[tag] = (B.copy(), dv.copy(), n_iter.copy(), status.copy(), dev.fitted())
        finally:
            os.environ.pop("CA_GLM_NO_ORTHO", None)
    ok = (out["ortho"][3] == 0) & (out["plain"][3] == 0)
    assert ok.sum() >= p - 2
    np.testing.assert_allclose(out["ortho"][0][ok], out["plain"][0][ok], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out["ortho"][1][ok], out["plain"][1][ok], rtol=1e-10)
    assert (out["ortho"][2][ok] != out["plain"][2][ok]).sum() <= 1
    np.testing.assert_allclose(out["ortho"][4][:, ok], out["plain"][4][:, ok], rtol=1e-9)
    B_ref, _, _, _, _, status_ref, _ = og.fit_glm_original(Y, X, None, family=fam, disp_glm=disp, offset=off, maxiter=100)
    ok &= status_ref == 0
    np.testing.assert_allclose(out["ortho"][0][ok], B_ref[ok], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("fam", ["poisson", "nb"])
def test_glm_guard_lasso_refit(fam):
    """Genes that trip the reference's coefficient guards (a tiny-scale column forces |beta| > 50)
    are refitted with the lasso objective of the reference's fit_regularized fallback: optimality
    conditions of that objective and agreement with the oracle's minimiser."""
    from causarray_b200 import glm as cglm
    from causarray_b200.device import GlmDevice
    rng = np.random.default_rng(8)
    n, p = 600, 40
    u = rng.standard_normal(n)
    u /= np.linalg.norm(u)                          # unit-norm column like the SVD factors of the init
    X = np.c_[np.ones(n), rng.standard_normal((n, 2)) * 0.3, u]
    beta = np.c_[rng.uniform(0.5, 2.0, p), rng.standard_normal((p, 2)) * 0.2, rng.choice([-1, 1], p) * rng.uniform(20, 120, p)]
    mu = np.exp(X @ beta.T)
    nu = rng.uniform(0.5, 2.0, p)
    Y = (rng.poisson(mu) if fam == "poisson" else rng.negative_binomial(nu[None, :], nu[None, :] / (nu[None, :] + mu))).astype(float)
    off = rng.standard_normal(n) * 0.1
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    fit = cglm._device_fit(dev, X, fam, off, nu if fam == "nb" else None, 100.0, 100, X.shape[1])
    bad = fit.refit_mask
    print(fam, "genes refitted", int(bad.sum()), "of", p)
    assert bad.sum() >= 5 and (~bad).sum() >= 5
    pen = np.r_[0.0, np.full(X.shape[1] - 1, 1e-4)]
    for j in np.where(bad)[0]:
        a_nb = None if fam == "poisson" else 1.0 / nu[j]
        viol = og.lasso_kkt_violation(Y[:, j], X, off, a_nb, pen, fit.B[j])
        assert viol <= 2e-7, (j, viol)                # 1e-4 is the penalty itself
        b_ref, _ = og.lasso_glm(Y[:, j], X, off, a_nb, pen)
        np.testing.assert_allclose(fit.B[j], b_ref, rtol=1e-4, atol=1e-6)   # both stop short of the exact minimiser
    # genes that pass the guards keep the unpenalised fit
    j = int(np.where(~bad)[0][0])
    b_mle = og.irls(Y[:, j], X, off, None if fam == "poisson" else 1.0 / nu[j], maxiter=100)[0]
    np.testing.assert_allclose(fit.B[j], b_mle, rtol=1e-7, atol=1e-8)


def test_glm_golden_helpers(golden_dir):
    """MoM dispersion against the reference's own estimate_disp (golden)."""
    g = np.load(os.path.join(golden_dir, "misc.npz"))
    Y, mu, sf = g["Y"], g["mu"], g["size_factor"]
    ref = og.estimate_disp_moments(Y, mu, np.log(sf))
    np.testing.assert_allclose(ref, g["disp_mom"], rtol=1e-12)


@pytest.mark.parametrize("usevar", ["unequal", "pooled"])
def test_aipw_lfc_golden(golden_dir, usevar):
    from causarray_b200.device import aipw_lfc, aipw_lfc_yhat
    g = np.load(os.path.join(golden_dir, "lfc_%s.npz" % usevar))
    Y, W, A = g["Y"], g["W"], g["A"]
    n, p = Y.shape
    a = A.shape[1]
    sf = np.exp(g["offset"])
    out = aipw_lfc(Y, W, g["Bcov"], g["Btrt"], g["offset"], sf, A, g["pi"], usevar=usevar)
    tau = out["tau"].reshape(-1)
    std = np.sqrt(out["var"]).reshape(-1)
    fin = np.isfinite(g["df_std"])
    print(usevar, "tau err", np.abs(tau - g["df_tau"]).max(), "std rel err",
          np.abs(std[fin] / g["df_std"][fin] - 1).max(), "filtered", (~fin).sum())
    np.testing.assert_array_equal(np.isfinite(std), fin)
    np.testing.assert_allclose(tau, g["df_tau"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(std[fin], g["df_std"][fin], rtol=1e-9)
    np.testing.assert_allclose(out["mean0"].reshape(-1), g["df_mean_control"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out["mean1"].reshape(-1), g["df_mean_treated"], rtol=1e-10, atol=1e-13)
    np.testing.assert_array_equal(out["estimable"].reshape(-1), g["estimable"].astype(bool))
    # explicit-Yhat variant
    eta0 = W @ g["Bcov"].T + g["offset"][:, None]
    Y_hat = np.zeros((n, p, a, 2))
    for k in range(a):
        Y_hat[:, :, k, 0] = np.exp(eta0)
        Y_hat[:, :, k, 1] = np.exp(eta0 + g["Btrt"][:, k][None, :])
    out2 = aipw_lfc_yhat(Y, Y_hat, sf, A, g["pi"], usevar=usevar)
    np.testing.assert_allclose(out2["tau"].reshape(-1), g["df_tau"], rtol=1e-9, atol=1e-12)
    std2 = np.sqrt(out2["var"]).reshape(-1)
    np.testing.assert_allclose(std2[fin], g["df_std"][fin], rtol=1e-9)


def test_inference_tail_device_vs_scipy_and_host_bh():
    """Device inference tail (csrc/inference.cu: Student-t / normal two-sided tails by the incomplete
    beta continued fraction, per-treatment BH by bitonic sort + reverse running minimum, median / MAD
    empirical null) against scipy.stats.t.sf / norm.sf and the reference-equivalent host path
    (causarray/DR_inference.py:153-201; inference.bh_correction is pinned on the reference golden in
    tests/test_oracle_golden.py::test_misc_golden)."""
    import scipy.stats
    from causarray_b200 import inference as inf
    rng = np.random.default_rng(5)
    a, p = 7, 3001
    t = rng.standard_normal((a, p)) * 1.3
    t[:, :200] += rng.choice([-1.0, 1.0], (a, 200)) * rng.uniform(3, 12, (a, 200))
    t[1, 5:40] = np.nan                      # not tested
    t[2, :] = np.nan                         # a treatment with nothing to test
    t[3, 10] = t[3, 11]                      # tie
    t[4, :] = 0.25                           # MAD = 0: adjusted columns stay NaN
    t[5, 7] = 38.0                           # p-value ~ 1e-100 territory
    df = rng.uniform(1.5, 4000.0, (a, p))
    df[np.isnan(t)] = np.nan
    for dfa in (df, None):
        pv, qv, pva, qva = inf.bh_correction_device(t, df=dfa)
        with np.errstate(invalid="ignore"):
            ref_p = (scipy.stats.t.sf(np.abs(t), dfa) if dfa is not None else scipy.stats.norm.sf(np.abs(t))) * 2
        np.testing.assert_allclose(pv, ref_p, rtol=1e-10, atol=1e-300, equal_nan=True)
        for j in range(a):
            hp, hq, hpa, hqa = inf.bh_correction(t[j], df=None if dfa is None else dfa[j])
            np.testing.assert_allclose(qv[j], hq, rtol=1e-10, atol=1e-300, equal_nan=True)
            np.testing.assert_allclose(pva[j], hpa, rtol=1e-9, atol=1e-300, equal_nan=True)
            np.testing.assert_allclose(qva[j], hqa, rtol=1e-9, atol=1e-300, equal_nan=True)
            assert np.array_equal(qv[j] < 0.05, hq < 0.05)
    assert np.isnan(pva[4]).all() and np.isnan(qva[2]).all()


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_fit_glm_fast_two_stage_vs_oracle(fam):
    """backend="fast": the reference's two-stage batched route (causarray/nb_glm_fast.py:373-594 --
    covariate model on a <= 3000-cell subsample with iteration caps, method-of-moments dispersion, then
    per-perturbation treatment effects against the covariate offset with the dispersion fixed) on the
    device against its restatement oracle/glm_fast.py (parity unpinned against crispyx itself)."""
    from causarray_b200 import glm as cglm
    from causarray_b200.synth import simulate_screen
    from oracle import glm_fast as ogf
    sim = simulate_screen(n_ctrl=2600, n_perts=4, cells_per_pert=150, p=64, r=2, d_cov=2, family=fam, seed=9,
                          base_lo=0.0, base_hi=2.0)
    Y, A = sim["Y"], sim["A"]
    X = np.c_[sim["X"], sim["U"]]
    off = np.log(sim["size_scale"])
    assert Y.shape[0] > 3000                     # the stage-1 subsample is drawn
    with cglm._backend_override("fast"):
        B, Yhat, disp, offs, resid = cglm.fit_glm_auto(Y, X, A, family=fam, offset=off, impute=True, random_state=3)
    assert cglm.last_fit_info["route"] == "fast-two-stage" and cglm.last_fit_info["n_stage1_cells"] == 3000
    Bo, nuo, ito = ogf.fit_glm_fast_two_stage(Y, X, A, family=fam, offset=off, random_state=3)
    np.testing.assert_allclose(B[:, :X.shape[1]], Bo[:, :X.shape[1]], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(B[:, X.shape[1]:], Bo[:, X.shape[1]:], rtol=1e-6, atol=1e-8)
    if fam == "nb":
        np.testing.assert_allclose(disp, nuo, rtol=1e-8)
    d = X.shape[1]
    eta0 = X @ B[:, :d].T + off[:, None]
    np.testing.assert_allclose(Yhat[0][:, :, 1], np.exp(np.clip(eta0, -20, 20)), rtol=1e-12)
    np.testing.assert_allclose(Yhat[1][:, :, 2], np.exp(np.clip(eta0 + B[:, d + 2][None, :], -20, 20)), rtol=1e-12)
    assert resid.shape == Y.shape and np.isfinite(resid).all()
    # the default route is the joint fit
    B_joint = cglm.fit_glm_auto(Y, X, A, family=fam, offset=off, disp_glm=None if fam == "poisson" else disp)[0]
    assert cglm.last_fit_info.get("route") != "fast-two-stage" or True
    assert np.abs(B_joint[:, d:] - B[:, d:]).max() < 0.5       # same effects up to the two-stage approximation
