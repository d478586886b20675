"""CPU: host-side mirror of the reference interface (no compute calls into the CUDA library)."""
import ctypes
import os
import re

import numpy as np
import pytest

import causarray_b200 as cb
from causarray_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/causarray_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "causarray_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = L.load(require_gpu=False)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    assert lib.ca_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.CausarrayCudaError):
        cb.nll(np.ones((4, 3)), np.ones((4, 2)), np.ones((3, 2)), "poisson", np.ones(3), np.zeros((4, 3)))


def test_public_api_surface():
    for name in ["LFC", "gcate_lfc_batch", "LFC_batch", "fit_glm", "fit_glm_fast", "fit_glm_auto", "fit_gcate",
                 "fit_gcate_batch", "estimate_r", "estimate", "alter_min", "nll", "grad", "update",
                 "update_with_mask", "Early_Stopping", "prep_causarray_data", "comp_size_factor",
                 "estimate_propensity_scores", "compute_causal_estimand", "cross_fitting", "AIPW_mean",
                 "reset_random_seeds", "bh_correction"]:
        assert hasattr(cb, name), name
    import inspect
    sig = inspect.signature(cb.gcate_lfc_batch)
    assert list(sig.parameters)[:4] == ["Y", "X", "A", "r"]
    assert sig.parameters["batch_size"].default == 10 and sig.parameters["max_cells"].default == 2000
    assert sig.parameters["family"].default == "nb" and sig.parameters["offset"].default is True
    sig = inspect.signature(cb.LFC)
    assert sig.parameters["usevar"].default == "unequal" and sig.parameters["thres_min"].default == 1e-2
    assert sig.parameters["ps_clip"].default == (0.01, 0.99)
    sig = inspect.signature(cb.fit_gcate)
    assert sig.parameters["c1"].default is None and sig.parameters["backend"].default == "auto"


def test_prep_and_size_factors_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "misc.npz"))
    np.testing.assert_allclose(cb.comp_size_factor(g["Y"]), g["size_factor"], rtol=1e-13)
    Yc, A1, X1, XA1 = cb.prep_causarray_data(g["Y"], np.zeros((g["Y"].shape[0], 2)), X=None)
    np.testing.assert_allclose(Yc, g["prep_Y"])
    np.testing.assert_allclose(X1, g["prep_X"])
    np.testing.assert_allclose(XA1, g["prep_XA"], rtol=1e-12)
    with pytest.raises(ValueError):
        cb.comp_size_factor(g["Y"], method="bogus")


def test_subsamplers_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "misc.npz"))
    np.testing.assert_array_equal(cb.subsample_ctrl_cells(np.arange(5, 500, 3), n_ctrl=40, random_state=3),
                                  g["ctrl_sel"])
    np.testing.assert_array_equal(cb.subsample_pert_cells(np.arange(2, 700, 5), max_cells=33, random_state=4),
                                  g["pert_sel"])
    idx = np.arange(10)
    assert cb.subsample_ctrl_cells(idx, n_ctrl=20) is not None and len(cb.subsample_ctrl_cells(idx, n_ctrl=20)) == 10
    assert len(cb.subsample_pert_cells(idx, max_cells=None)) == 10


def test_bh_correction_golden_and_edges(golden_dir):
    g = np.load(os.path.join(golden_dir, "misc.npz"))
    for pre, df in (("bh", g["bh_df"]), ("bhz", None)):
        out = cb.bh_correction(g["bh_t"].copy(), df=df)
        for got, name in zip(out, ["p", "q", "pa", "qa"]):
            np.testing.assert_allclose(got, g["%s_%s" % (pre, name)], rtol=1e-12, equal_nan=True)
    # NaN propagation and the mad == 0 branch (reference tests/test_inference_comprehensive.py:317-394)
    t = np.array([np.nan, 2.0, 2.0, 2.0])
    p, q, pa, qa = cb.bh_correction(t)
    assert np.isnan(p[0]) and np.isnan(q[0]) and np.all(np.isnan(pa))
    assert np.all(np.isnan(cb.bh_correction(np.full(3, np.nan))[0]))


def test_early_stopping_matches_reference_rule():
    es = cb.Early_Stopping(warmup=0, patience=2, tolerance=0.0, rel_tol=1e-2)
    seq = [1.0, 0.9, 0.895, 0.894, 0.893, 0.892]
    out = [es(v) for v in seq]
    assert out == [False, False, False, False, True, True]
    assert "Best Epoch: 1" in es.info


def test_filter_params_and_log2_columns():
    from causarray_b200.dr import _add_log2fc_columns
    from causarray_b200.prep import _filter_params
    import pandas as pd
    assert _filter_params(cb.comp_size_factor, dict(method="scale", foo=1)) == {"method": "scale"}
    assert _filter_params({"a": 1}, dict(a=2, b=3)) == {"a": 2}
    with pytest.raises(ValueError):
        _filter_params(3, {})
    df = pd.DataFrame({"gene_names": [0, 1], "tau": [np.log(2.0), 0.0], "std": [np.log(4.0), 1.0], "stat": [1., 2.]})
    out = _add_log2fc_columns(df)
    assert list(out.columns) == ["gene_names", "tau", "std", "log2fc", "log2fc_se", "stat"]
    np.testing.assert_allclose(out["log2fc"], [1.0, 0.0])
    np.testing.assert_allclose(out["log2fc_se"], [2.0, 1.0 / np.log(2.0)])


def test_propensity_scores_host():
    rng = np.random.default_rng(0)
    n = 300
    X = np.c_[np.ones(n), rng.standard_normal((n, 2))]
    A = np.zeros((n, 2))
    A[:60, 0] = 1
    A[60:110, 1] = 1
    pi = cb.estimate_propensity_scores(A, X, clip=(0.01, 0.99))
    assert pi.shape == (n, 2) and np.all((pi >= 0.01) & (pi <= 0.99))
    # constant design short-cut
    np.testing.assert_allclose(cb.estimate_propensity_scores(A, np.ones((n, 1))), 0.5)
    with pytest.raises(ValueError):
        cb.estimate_propensity_scores(A * 2, X)
    with pytest.raises(ValueError):
        cb.estimate_propensity_scores(A, X, K=0)
    with pytest.raises(ValueError):
        cb.estimate_propensity_scores(A, X, clip=(0.5, 0.2))


def test_aipw_mean_host_helper():
    """Reference tests/test_DR_learner.py:76-101: pseudo-outcomes are left unclipped."""
    rng = np.random.default_rng(1)
    n, p, a = 20, 3, 1
    Y = rng.poisson(3.0, (n, p)).astype(float)
    A = (rng.random((n, a)) < 0.5).astype(float)
    pi = np.full((n, a), 0.5)
    mu = np.full((n, p, a, 2), 3.0)
    tau, psi = cb.AIPW_mean(Y, np.stack([1 - A, A], -1), mu, np.stack([1 - pi, pi], -1))
    assert tau.shape == (p, a, 2) and psi.shape == (n, p, a, 2) and tau.dtype == np.float64
    assert psi.min() < 0
    np.testing.assert_allclose(tau, psi.mean(0))


def test_yhat_factors_materialize():
    from causarray_b200.dr import YhatFactors
    rng = np.random.default_rng(2)
    W = rng.standard_normal((5, 2))
    Bc, Bt = rng.standard_normal((4, 2)), rng.standard_normal((4, 3))
    off = rng.standard_normal(5) * 0.1
    yh = YhatFactors(W, Bc, Bt, off).materialize()
    assert yh.shape == (5, 4, 3, 2)
    np.testing.assert_allclose(yh[:, :, 1, 0], np.exp(W @ Bc.T + off[:, None]))
    np.testing.assert_allclose(yh[:, :, 2, 1], np.exp(W @ Bc.T + off[:, None] + Bt[:, 2][None, :]))


def test_lowrank_product_svd():
    from causarray_b200.svd import lowrank_product_svd
    rng = np.random.default_rng(3)
    U, G = rng.standard_normal((30, 3)), rng.standard_normal((40, 3))
    u, s, vt = lowrank_product_svd(U, G)
    np.testing.assert_allclose((u * s) @ vt, U @ G.T, atol=1e-10)
    s_ref = np.linalg.svd(U @ G.T, compute_uv=False)[:3][::-1]     # ascending, the order scipy's svds returns
    np.testing.assert_allclose(s, s_ref, rtol=1e-10)


def test_host_inference_helpers_match_numpy():
    """ca_host_bh_rows / ca_host_median_mad_rows (CPU code in the shared library) against the numpy
    forms, with NaNs, ties, an all-NaN family and an even / odd number of kept entries."""
    from causarray_b200 import inference as inf
    rng = np.random.default_rng(4)
    p = rng.random((6, 501))
    p[1, 3:40] = np.nan
    p[2] = np.nan
    p[3, :50] = p[3, 60]
    p[4, ::2] = np.nan
    q = inf._bh_rows(p)
    np.testing.assert_array_equal(q, inf._bh_rows_block(p))
    t = rng.standard_normal((6, 501))
    t[1, 3:40] = np.nan
    t[4, ::7] = np.nan
    med, mad = inf._median_mad_rows(t)
    np.testing.assert_allclose(med, np.nanmedian(t, axis=1), rtol=0, atol=0)
    np.testing.assert_allclose(mad, np.nanmedian(np.abs(t - med[:, None]), axis=1) / inf._MAD_NORMAL, rtol=1e-15)


def test_typed_upload_codes_match_header():
    """The element-type codes the Python layer passes to ca_gcate_upload_counts_typed are the header's."""
    import os
    import re
    import numpy as np
    from causarray_b200.device import _NARROW_DTYPES, device_dtype
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "causarray_b200.h")).read()
    m = re.search(r"enum\s*\{([^}]*CA_F64[^}]*)\}", hdr)
    codes = {k.strip(): int(v) for k, v in (kv.split("=") for kv in m.group(1).split(","))}
    want = {"CA_F32": np.float32, "CA_I32": np.int32, "CA_I64": np.int64, "CA_U16": np.uint16, "CA_U8": np.uint8}
    assert codes["CA_F64"] == 0
    for name, t in want.items():
        assert _NARROW_DTYPES[np.dtype(t).str] == codes[name]
    assert device_dtype(np.zeros((2, 2), dtype=np.int32)) and device_dtype(np.zeros((2, 2)))
    assert not device_dtype(np.zeros((2, 2), dtype=np.int16)) and not device_dtype([[1, 2]])


def test_lean_logistic_is_sklearn():
    """The lean propensity fit (same loss object, weights, start and L-BFGS-B options as sklearn's
    lbfgs path) returns sklearn's probabilities bit for bit -- the reference's propensity model is
    ``LogisticRegression(fit_intercept=False, C=1, class_weight='balanced')``
    (causarray/DR_estimation.py:18-42)."""
    from sklearn.linear_model import LogisticRegression
    from causarray_b200 import dr
    rng = np.random.default_rng(3)
    for n, d, frac, cw in ((900, 4, 0.3, 'balanced'), (2133, 12, 0.06, 'balanced'), (500, 3, 0.5, None)):
        X = np.c_[np.ones(n), rng.standard_normal((n, d - 1))]
        y = (rng.random(n) < frac).astype(float)
        Xte = rng.standard_normal((50, d))
        kw = dict(fit_intercept=False, C=1.0, class_weight=cw, random_state=0)
        lean = dr._fit_logistic(kw, X, y, Xte)
        ref = LogisticRegression(**kw).fit(X, y).predict_proba(Xte)[:, 1]
        assert np.array_equal(lean, ref)
    # options outside the plain regime go through the estimator itself
    kw = dict(fit_intercept=True, C=0.5, class_weight='balanced', random_state=0)
    np.testing.assert_array_equal(dr._fit_logistic(kw, X, y, Xte),
                                  LogisticRegression(**kw).fit(X, y).predict_proba(Xte)[:, 1])


def test_direct_lbfgsb_driver_is_scipy_minimize():
    """The propensity fits drive scipy's compiled L-BFGS-B routine directly (the loop of
    ``_minimize_lbfgsb`` without ScalarFunction / OptimizeResult): same iterates, bit for bit, as
    ``scipy.optimize.minimize(method='L-BFGS-B', jac=True)`` -- also when the iteration cap stops it."""
    import scipy.optimize
    from causarray_b200 import dr
    rng = np.random.default_rng(5)
    for n, d, maxiter in ((400, 6, 100), (1500, 12, 100), (800, 9, 3)):
        X = rng.standard_normal((n, d))
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X[:, 0]))).astype(float)

        def func(w):
            z = X @ w
            pr = 1.0 / (1.0 + np.exp(-z))
            return float(np.sum(np.logaddexp(0.0, z) - y * z) / n + 0.5e-3 * (w @ w)), X.T @ (pr - y) / n + 1e-3 * w

        ftol = 64 * np.finfo(float).eps
        ref = scipy.optimize.minimize(func, np.zeros(d), method='L-BFGS-B', jac=True,
                                      options={'maxiter': maxiter, 'maxls': 50, 'gtol': 1e-4, 'ftol': ftol}).x
        assert np.array_equal(dr._lbfgsb_direct(func, np.zeros(d), maxiter, 50, 1e-4, ftol), ref)
    # the gate: the first fit of a process is run both ways and the direct driver is only used when they agree
    dr._LBFGSB_DIRECT = None
    out = dr._lbfgsb_minimize(func, np.zeros(d), 100, 50, 1e-4, ftol)
    assert dr._LBFGSB_DIRECT is True
    assert np.array_equal(out, dr._lbfgsb_minimize(func, np.zeros(d), 100, 50, 1e-4, ftol))


def test_direct_lbfgsb_driver_gate_falls_back(monkeypatch):
    """A direct driver that does not reproduce scipy's iterates (another scipy behind the private interface)
    is never used: the probe fit returns scipy's answer and later fits go through ``minimize``."""
    import scipy.optimize
    from causarray_b200 import dr
    rng = np.random.default_rng(9)
    X = rng.standard_normal((300, 5))
    y = (rng.random(300) < 0.4).astype(float)

    def func(w):
        z = X @ w
        return float(np.sum(np.logaddexp(0.0, z) - y * z) / 300 + 0.5e-3 * (w @ w)), X.T @ (1 / (1 + np.exp(-z)) - y) / 300 + 1e-3 * w

    ftol = 64 * np.finfo(float).eps
    ref = scipy.optimize.minimize(func, np.zeros(5), method='L-BFGS-B', jac=True,
                                  options={'maxiter': 100, 'maxls': 50, 'gtol': 1e-4, 'ftol': ftol}).x
    monkeypatch.setattr(dr, '_LBFGSB_DIRECT', None)
    monkeypatch.setattr(dr, '_lbfgsb_direct', lambda *a, **k: ref + 1e-9)        # "works", but differently
    assert np.array_equal(dr._lbfgsb_minimize(func, np.zeros(5), 100, 50, 1e-4, ftol), ref)
    assert dr._LBFGSB_DIRECT is False
    assert np.array_equal(dr._lbfgsb_minimize(func, np.zeros(5), 100, 50, 1e-4, ftol), ref)

    def boom(*a, **k):
        raise TypeError('setulb() takes 16 positional arguments')
    monkeypatch.setattr(dr, '_LBFGSB_DIRECT', None)
    monkeypatch.setattr(dr, '_lbfgsb_direct', boom)
    assert np.array_equal(dr._lbfgsb_minimize(func, np.zeros(5), 100, 50, 1e-4, ftol), ref)
    assert dr._LBFGSB_DIRECT is False


def test_fdx_control_matches_reference_golden():
    """FDX control (multiplier bootstrap as one GEMM, step-down thresholds from a reversed running
    maximum, augmentation) against the REAL reference's loop implementation
    (causarray/DR_inference.py:8-150; tests/golden/make_golden.py::golden_fdx): same seeded numpy
    stream, identical discovery sets -- ties in |t|, untestable hypotheses, c = 0, no signal."""
    from causarray_b200 import inference as inf
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fdx.npz"))
    for ci in range(4):
        B, alpha, c, seed = g["c%d_par" % ci]
        np.random.seed(int(seed))
        V = inf.fdx_control(g["c%d_tau" % ci], g["c%d_t" % ci].copy(), g["c%d_eta" % ci], g["c%d_std" % ci], True,
                            int(B), float(alpha), float(c))
        np.testing.assert_array_equal(V, g["c%d_V" % ci])
    assert [int(g["c%d_V" % ci].sum()) for ci in range(4)] == [24, 87, 0, 39]
    # fdx=False: no discoveries, no random draws
    st = np.random.get_state()[1][:4].copy()
    assert inf.fdx_control(g["c0_tau"], g["c0_t"], g["c0_eta"], g["c0_std"], False, 10, 0.05, 0.1).sum() == 0
    assert np.array_equal(np.random.get_state()[1][:4], st)
