"""End-to-end GPU parity: fit_gcate + LFC through the public API against the golden run of the
REAL reference orchestration (tests/golden/make_golden.py::golden_pipeline; the statsmodels
calls of the reference were served by oracle/glm.py, so the GLM piece is 'parity unpinned')."""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6   # north-star tolerance on LFC estimates and standard errors (fp64)


def _subspace_gap(U1, U2):
    q1, q2 = np.linalg.qr(U1)[0], np.linalg.qr(U2)[0]
    return np.linalg.norm(q1 @ q1.T - q2 @ q2.T)


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_fit_gcate_lfc_golden(golden_dir, fam):
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "pipeline_%s.npz" % fam))
    Y, X, A, disp = g["Y"], g["X"], g["A"], g["disp"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res1, res2 = cb.fit_gcate(Y.copy(), X, A, 2, family=fam, disp_glm=disp.copy(),
                                  kwargs_es_1=dict(max_iters=30), kwargs_es_2=dict(max_iters=30))
        ws = res2.pop("_ws")
        d = X.shape[1] + A.shape[1]
        print(fam, "n_iter", res1["n_iter"], res2["n_iter"], "hist1 end", res1["hist"][-1], g["s1_hist"][-1],
              "len", len(res1["hist"]), len(g["s1_hist"]))
        print(" size factor diff", np.abs(res2["kwargs_glm"]["size_factor"] - g["size_factor"]).max())
        print(" stage1 |dA|", np.abs(res1["X_U"] - g["s1_A"]).max(), "|dB|", np.abs(res1["B_Gamma"] - g["s1_B"]).max(),
              "subspace gap", _subspace_gap(res1["X_U"][:, d:], g["s1_A"][:, d:]))
        print(" stage2 |dA|", np.abs(res2["X_U"] - g["s2_A"]).max(), "|dB|", np.abs(res2["B_Gamma"] - g["s2_B"]).max())
        W = np.c_[X, res2["U"]]
        off = np.log(res2["kwargs_glm"]["size_factor"])
        df, est = cb.LFC(Y.copy(), W, A, W_A=W, family=fam, offset=off, disp_glm=disp.copy(), backend="original",
                         _glm=ws.glm)
        ws.close()
    np.testing.assert_allclose(res2["kwargs_glm"]["size_factor"], g["size_factor"], rtol=1e-12)
    assert len(res1["hist"]) == len(g["s1_hist"])
    np.testing.assert_allclose(res1["hist"], g["s1_hist"], rtol=1e-7)
    np.testing.assert_allclose(res2["hist"], g["s2_hist"], rtol=1e-7)
    np.testing.assert_allclose(est["pi_hat"], g["pi_hat"], rtol=1e-6)
    tau, std = df["tau"].to_numpy(), df["std"].to_numpy()
    fin = np.isfinite(g["df_std"])
    np.testing.assert_array_equal(np.isfinite(std), fin)
    err_tau = np.abs(tau - g["df_tau"]) / np.maximum(np.abs(g["df_tau"]), 1e-3)
    err_std = np.abs(std[fin] / g["df_std"][fin] - 1)
    print(" LFC max rel err tau", err_tau.max(), "std", err_std.max())
    assert err_tau.max() <= REL_TOL
    assert err_std.max() <= REL_TOL
    # p-values and BH rejections at alpha = 0.05 unchanged
    rej = df["padj"].to_numpy() < 0.05
    rej_ref = g["df_padj"] < 0.05
    np.testing.assert_array_equal(rej, rej_ref)
    np.testing.assert_allclose(df["pvalue"].to_numpy()[fin], g["df_pvalue"][fin], rtol=1e-4, atol=1e-12)
    assert list(df.columns) == ['gene_names', 'tau', 'std', 'log2fc', 'log2fc_se', 'stat', 'rej', 'pvalue', 'padj',
                                'pvalue_emp_null_adj', 'padj_emp_null_adj', 'mean_control', 'mean_treated',
                                'estimable', 'trt']


def test_gcate_lfc_batch_smoke():
    """Batch driver: columns, coverage of every perturbation, determinism (reference
    tests/test_batch_fitting.py:294-326)."""
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=150, n_perts=6, cells_per_pert=40, p=150, r=2, seed=3, base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    kw = dict(batch_size=2, max_cells=2000, n_ctrl=100, family="nb",
              gcate_kwargs=dict(disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=8), kwargs_es_2=dict(max_iters=8)),
              lfc_kwargs=dict(usevar="unequal", disp_glm=sim["disp"]), random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = cb.gcate_lfc_batch(Y, X, A, 2, **kw)
        df2 = cb.gcate_lfc_batch(Y, X, A, 2, **kw)
    assert len(df) == Y.shape[1] * A.shape[1]
    assert sorted(df["trt"].unique()) == list(range(6))
    assert sorted(df["batch"].unique()) == [0, 1, 2]
    for c in ["tau", "std", "log2fc", "pvalue", "padj", "batch", "trt", "mean_control", "estimable"]:
        assert c in df.columns
    np.testing.assert_allclose(df["log2fc"], df["tau"] / np.log(2.0))
    np.testing.assert_array_equal(df["tau"].to_numpy(), df2["tau"].to_numpy())


def test_gcate_lfc_batch_cache_round_trip(tmp_path):
    """Resumable HDF5 cache of the batch driver (reference DR_learner.py:705-760): a second call with the
    same cache_path reads every batch back and returns the same table.  Needs pytables (not in every image)."""
    pytest.importorskip("tables")
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=120, n_perts=4, cells_per_pert=40, p=120, r=2, seed=5, base_lo=0.0, base_hi=2.0)
    kw = dict(batch_size=2, max_cells=2000, n_ctrl=100, family="nb",
              gcate_kwargs=dict(disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=6), kwargs_es_2=dict(max_iters=6)),
              lfc_kwargs=dict(usevar="unequal", disp_glm=sim["disp"]), random_state=0)
    path = str(tmp_path / "lfc_cache.h5")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = cb.gcate_lfc_batch(sim["Y"], sim["X"], sim["A"], 2, cache_path=path, **kw)
        df2 = cb.gcate_lfc_batch(sim["Y"], sim["X"], sim["A"], 2, cache_path=path, **kw)
    assert os.path.exists(path)
    np.testing.assert_array_equal(df["tau"].to_numpy(), df2["tau"].to_numpy())
    np.testing.assert_array_equal(df["batch"].to_numpy(), df2["batch"].to_numpy())


def test_estimate_r_golden(golden_dir):
    """JIC sweep over the number of latent factors against the REAL reference's estimate_r
    (causarray/gcate.py:196-325; golden generated by tests/golden/make_golden.py)."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "estimate_r.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = cb.estimate_r(g["Y"].copy(), g["X"], g["A"], 3, family="nb", disp_glm=g["disp"].copy(),
                           kwargs_es_1=dict(max_iters=30), kwargs_es_2=dict(max_iters=30))
    assert list(df.columns) == ["r", "deviance", "nu", "JIC"]
    np.testing.assert_array_equal(df["r"].to_numpy(), g["r"])
    print(df.to_string())
    np.testing.assert_allclose(df["nu"].to_numpy(), g["nu"], rtol=1e-12)
    np.testing.assert_allclose(df["deviance"].to_numpy(), g["deviance"], rtol=1e-6)
    np.testing.assert_allclose(df["JIC"].to_numpy(), g["JIC"], rtol=1e-6)


def test_gcate_lfc_batch_golden(golden_dir):
    """Batch driver against the REAL reference's gcate_lfc_batch (control subsampling, batch
    partition, per-batch GCATE + LFC, frame layout; causarray/DR_learner.py:584-866)."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "batch.npz"))
    kw = dict(batch_size=2, max_cells=2000, n_ctrl=120, family="nb",
              gcate_kwargs=dict(disp_glm=g["disp"].copy(), kwargs_es_1=dict(max_iters=20), kwargs_es_2=dict(max_iters=20)),
              lfc_kwargs=dict(usevar="unequal", disp_glm=g["disp"].copy()), random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = cb.gcate_lfc_batch(g["Y"].copy(), g["X"], g["A"], 2, **kw)
    assert list(df.columns) == list(g["columns"])
    assert len(df) == len(g["df_tau"])
    np.testing.assert_array_equal(df["batch"].to_numpy().astype(float), g["df_batch"])
    for c in ("trt", "gene_names"):       # labels: numeric in the reference's frame when no names are passed
        want = g["df_" + c]
        got = df[c].to_numpy()
        np.testing.assert_array_equal(got.astype(str) if want.dtype.kind in "US" else got.astype(float), want)
    np.testing.assert_array_equal(df["estimable"].to_numpy().astype(float), g["df_estimable"])
    worst = {}
    for c in ["tau", "std", "log2fc", "log2fc_se", "stat", "mean_control", "mean_treated"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), c
        worst[c] = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3)))
        assert worst[c] <= 1e-6, (c, worst[c])
    for c in ["pvalue", "padj", "pvalue_emp_null_adj", "padj_emp_null_adj"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12, equal_nan=True)
        assert np.array_equal(a < 0.05, b < 0.05), c
    print("batch golden worst relative errors", worst)


@pytest.mark.parametrize("tag,cross_est", [("k1", False), ("k2", True)])
def test_lfc_with_own_nuisance_fits_golden(golden_dir, tag, cross_est):
    """LFC fitting its own outcome GLM and logistic propensities, in-sample and with two-fold
    cross-fitting (causarray/DR_estimation.py:472-618), against the REAL reference run with
    backend="original" (its statsmodels route, whose two-fold imputation quirk is mirrored under the
    same flag)."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "lfc_fit_%s.npz" % tag))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df, est = cb.LFC(g["Y"].copy(), g["W"], g["A"], W_A=g["W"], family="nb", offset=g["offset"],
                         disp_glm=g["disp"].copy(), cross_est=cross_est, usevar="unequal", backend="original")
    assert list(df.columns) == list(g["columns"])
    np.testing.assert_allclose(est["pi_hat"], g["pi_hat"], rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(df["estimable"].to_numpy().astype(np.uint8), g["estimable"])
    for c in ["tau", "std", "stat", "mean_control", "mean_treated"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), c
        err = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3)))
        print(tag, c, "max rel err", err)
        assert err <= 1e-6, (c, err)
    for c in ["pvalue", "padj"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12, equal_nan=True)
        assert np.array_equal(a < 0.05, b < 0.05), c


@pytest.mark.parametrize("tag", ["k2_odd", "k3"])
def test_lfc_cross_fitted_held_out_prediction_golden(golden_dir, tag):
    """Default route of cross-fitting: the outcome model of each training fold predicts the HELD-OUT
    cells with their own offsets (reference: DR_estimation.py:590-613 with nb_glm_fast.py:273-297), for
    fold shapes the statsmodels route cannot run (odd n with K = 2; K = 3).  Golden: fold-wise IRLS
    restatement for Y_hat, then the REAL reference's LFC (cross-fitted propensities, AIPW, BH)."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "lfc_crossfit_%s.npz" % tag))
    K = int(g["K"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df, est = cb.LFC(g["Y"].copy(), g["W"], g["A"], W_A=g["W"], family="nb", offset=g["offset"],
                         disp_glm=g["disp"].copy(), cross_est=True, K=K, usevar="unequal")
    np.testing.assert_allclose(est["pi_hat"], g["pi_hat"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.asarray(est["Y_hat"], dtype=float), g["Y_hat"], rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(df["estimable"].to_numpy().astype(np.uint8), g["estimable"])
    for c in ["tau", "std", "stat", "mean_control", "mean_treated"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), c
        err = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3)))
        assert err <= 1e-6, (c, err)
    for c in ["pvalue", "padj"]:
        a, b = df[c].to_numpy().astype(float), g["df_" + c]
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12, equal_nan=True)
        assert np.array_equal(a < 0.05, b < 0.05), c


def test_fit_gcate_batch_warm_start_and_fit_glm_impute_golden(golden_dir):
    """fit_gcate_batch (control / perturbed-cell subsampling, warm start of U across batches,
    causarray/gcate.py:328-609) and fit_glm with its own dispersion estimate and imputation
    (causarray/gcate_glm.py:95-272) against the REAL reference."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "more.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = cb.fit_gcate_batch(g["Y"].copy(), g["X"], g["A"], 2, batch_size=2, max_cells=60, n_ctrl=100, family="nb",
                                 disp_glm=g["disp"].copy(), warm_start_U=True, random_state=3,
                                 kwargs_es_1=dict(max_iters=15), kwargs_es_2=dict(max_iters=15))
    assert len(res) == int(g["n_batches"])
    for i, b in enumerate(res):
        np.testing.assert_array_equal(np.asarray(b["cell_idx"]), g["b%d_cell_idx" % i])
        np.testing.assert_array_equal(np.asarray(b["ctrl_idx"]), g["b%d_ctrl_idx" % i])
        np.testing.assert_array_equal(np.asarray(b["pert_idx"]), g["b%d_pert_idx" % i])
        assert len(b["res_1"]["hist"]) == len(g["b%d_hist1" % i]) and len(b["res_2"]["hist"]) == len(g["b%d_hist2" % i])
        np.testing.assert_allclose(b["res_2"]["hist"], g["b%d_hist2" % i], rtol=1e-8)
        np.testing.assert_allclose(b["res_2"]["X_U"], g["b%d_X_U" % i], atol=2e-6)
        np.testing.assert_allclose(b["res_2"]["B_Gamma"], g["b%d_B_Gamma" % i], atol=2e-6)
        ws = b["res_2"].pop("_ws", None)
        if ws is not None:
            ws.close()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        B, Yhat, disp, offsets, resid = cb.fit_glm(g["Y"].copy(), g["X"], g["A"], family="nb", disp_glm=None, offset=True,
                                                   impute=True, maxiter=100)
    np.testing.assert_allclose(offsets, g["glm_offsets"], rtol=1e-12)
    np.testing.assert_allclose(disp, g["glm_disp"], rtol=1e-8)
    np.testing.assert_allclose(B, g["glm_B"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(Yhat[0], g["glm_Yhat0"], rtol=1e-7)
    np.testing.assert_allclose(Yhat[1], g["glm_Yhat1"], rtol=1e-7)
    np.testing.assert_allclose(resid, g["glm_resid"], rtol=1e-6, atol=1e-7)


def test_lfc_option_variants_golden(golden_dir):
    """LFC option variants against the REAL reference: pooled variance + cell mask + unclipped
    propensities + non-default filters (NB with the dispersion estimated inside), and Poisson."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "lfc_options.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df1, est1 = cb.LFC(g["Y"].copy(), g["W"], g["A"], W_A=g["W"], family="nb", offset=True, mask=g["mask"].copy(),
                           usevar="pooled", ps_clip=None, thres_min=0.05, thres_diff=0.02)
        df2, est2 = cb.LFC(g["Y"].copy(), g["W"], g["A"], W_A=g["W"], family="poisson", offset=True, usevar="unequal")
    for tag, df, est in (("v1", df1, est1), ("v2", df2, est2)):
        np.testing.assert_allclose(est["pi_hat"], g[tag + "_pi_hat"], rtol=1e-9, atol=1e-12)
        np.testing.assert_array_equal(df["estimable"].to_numpy().astype(np.uint8), g[tag + "_estimable"])
        for c in ["tau", "std", "stat", "mean_control", "mean_treated"]:
            a, b = df[c].to_numpy().astype(float), g[tag + "_" + c]
            fin = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), fin), (tag, c)
            err = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3)))
            print(tag, c, "max rel err", err)
            assert err <= 1e-6, (tag, c, err)
        for c in ["pvalue", "padj"]:
            a, b = df[c].to_numpy().astype(float), g[tag + "_" + c]
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-12, equal_nan=True)
            assert np.array_equal(a < 0.05, b < 0.05), (tag, c)


def test_fit_gcate_option_variants_golden(golden_dir):
    """fit_gcate with no offset, an explicit c1 and non-default line-search / early-stopping settings
    against the REAL reference."""
    import causarray_b200 as cb
    g = np.load(os.path.join(golden_dir, "gcate_options.npz"))
    kw = dict(family="nb", disp_glm=g["disp"].copy(), offset=False, c1=0.2,
              kwargs_ls_1=dict(alpha=0.05, beta=0.6, tol=1e-3), kwargs_ls_2=dict(alpha=0.2, max_iters=15),
              kwargs_es_1=dict(max_iters=25, patience=3, rel_tol=1e-5), kwargs_es_2=dict(max_iters=12, tolerance=1e-6))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res1, res2 = cb.fit_gcate(g["Y"].copy(), g["X"], g["A"], 3, **kw)
    res2.pop("_ws").close()
    np.testing.assert_allclose(np.ravel(res2["kwargs_glm"]["size_factor"]), g["size_factor"], rtol=1e-12)
    np.testing.assert_allclose(np.ravel(res2["kwargs_glm"]["disp_glm"]), g["disp"], rtol=1e-8)
    assert len(res1["hist"]) == len(g["s1_hist"]) and len(res2["hist"]) == len(g["s2_hist"])
    np.testing.assert_allclose(res1["hist"], g["s1_hist"], rtol=1e-8)
    np.testing.assert_allclose(res2["hist"], g["s2_hist"], rtol=1e-8)
    np.testing.assert_allclose(res1["X_U"], g["s1_A"], atol=2e-6)
    np.testing.assert_allclose(res1["B_Gamma"], g["s1_B"], atol=2e-6)
    np.testing.assert_allclose(res2["X_U"], g["s2_A"], atol=2e-6)
    np.testing.assert_allclose(res2["B_Gamma"], g["s2_B"], atol=2e-6)


def test_gcate_lfc_batch_sparse_input_matches_dense():
    """A scipy.sparse count matrix (kept sparse, densified one batch at a time) gives the dense result."""
    import scipy.sparse as sp
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=120, n_perts=4, cells_per_pert=30, p=90, r=2, seed=5, base_lo=-0.5, base_hi=1.5)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    kw = dict(batch_size=2, max_cells=2000, n_ctrl=100, family="nb",
              gcate_kwargs=dict(kwargs_es_1=dict(max_iters=6), kwargs_es_2=dict(max_iters=6)),
              lfc_kwargs=dict(usevar="unequal"), random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_d = cb.gcate_lfc_batch(Y, X, A, 2, **kw)
        df_s = cb.gcate_lfc_batch(sp.csr_matrix(Y), X, A, 2, **kw)
        df_c = cb.gcate_lfc_batch(sp.csc_matrix(Y), X, A, 2, **kw)
    for c in ["tau", "std", "pvalue", "padj"]:
        np.testing.assert_array_equal(df_d[c].to_numpy(), df_s[c].to_numpy())
        np.testing.assert_array_equal(df_d[c].to_numpy(), df_c[c].to_numpy())


def test_fit_gcate_and_lfc_accept_sparse_counts():
    """fit_gcate + LFC on a scipy.sparse count matrix (CSR upload, expanded on the device; also a
    non-canonical COO-built matrix with duplicate entries) reproduce the dense results bit for bit."""
    import scipy.sparse as sp
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=110, n_perts=2, cells_per_pert=35, p=75, r=2, seed=9, base_lo=-0.5, base_hi=1.5)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    kw = dict(family="nb", disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=6), kwargs_es_2=dict(max_iters=6))

    def run(Yin):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r1, r2 = cb.fit_gcate(Yin, X, A, 2, **kw)
            ws = r2.pop("_ws")
            W = np.c_[X, r2["U"]]
            df, _ = cb.LFC(Yin, W, A, W_A=W, family="nb", offset=np.log(r2["kwargs_glm"]["size_factor"]),
                           disp_glm=sim["disp"], _glm=ws.glm, _trusted=True)
            ws.close()
        return r2["B_Gamma"], df

    B_d, df_d = run(Y)
    B_s, df_s = run(sp.csr_matrix(Y))
    np.testing.assert_array_equal(B_d, B_s)
    np.testing.assert_array_equal(df_d["tau"].to_numpy(), df_s["tau"].to_numpy())
    # duplicates that sum to the count, unsorted: y = (y - 1) + 1 for every y >= 2
    i, j = np.nonzero(Y)
    v = Y[i, j]
    big = v >= 2
    coo = sp.coo_matrix((np.r_[np.where(big, v - 1, v), np.ones(big.sum())], (np.r_[i, i[big]], np.r_[j, j[big]])),
                        shape=Y.shape)
    B_c, df_c = run(coo.tocsr())
    np.testing.assert_array_equal(B_d, B_c)
    np.testing.assert_array_equal(df_d["std"].to_numpy(), df_c["std"].to_numpy())


def test_integer_and_float32_counts_match_float64():
    """Count matrices of the usual storage types (int32 / int64 / float32 / uint16 / uint8) are uploaded
    as they are and widened on the device: fit_gcate + LFC and the batch driver reproduce the float64
    results bit for bit, and the device-side input checks see the same matrix."""
    import causarray_b200 as cb
    from causarray_b200.device import GcateDevice
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=110, n_perts=2, cells_per_pert=35, p=75, r=2, seed=9, base_lo=-0.5, base_hi=1.5)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    assert Y.max() < 65536
    kw = dict(family="nb", disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=6), kwargs_es_2=dict(max_iters=6))

    def run(Yin):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r1, r2 = cb.fit_gcate(Yin, X, A, 2, **kw)
            ws = r2.pop("_ws")
            W = np.c_[X, r2["U"]]
            df, _ = cb.LFC(Yin, W, A, W_A=W, family="nb", offset=np.log(r2["kwargs_glm"]["size_factor"]),
                           disp_glm=sim["disp"], _glm=ws.glm, _trusted=True)
            ws.close()
        return r2["B_Gamma"], df

    B_d, df_d = run(Y)
    types = [np.int32, np.int64, np.float32, np.uint16] + ([np.uint8] if Y.max() < 256 else [])
    for t in types:
        B_t, df_t = run(np.ascontiguousarray(Y.astype(t)))
        np.testing.assert_array_equal(B_d, B_t)
        for c in ["tau", "std", "pvalue", "padj"]:
            np.testing.assert_array_equal(df_d[c].to_numpy(), df_t[c].to_numpy())
    # the batch driver slices rows of the integer matrix and uploads the slices
    bkw = dict(batch_size=1, max_cells=2000, n_ctrl=100, family="nb",
               gcate_kwargs=dict(kwargs_es_1=dict(max_iters=5), kwargs_es_2=dict(max_iters=5)), random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_f = cb.gcate_lfc_batch(Y, X, A, 2, **bkw)
        df_i = cb.gcate_lfc_batch(Y.astype(np.int32), X, A, 2, **bkw)
    for c in ["tau", "std", "pvalue", "padj"]:
        np.testing.assert_array_equal(df_f[c].to_numpy(), df_i[c].to_numpy())
    # input checks on the widened copy: a NaN and a fractional entry in float32, an empty gene in uint8
    n, p = Y.shape
    Yf = Y.astype(np.float32)
    Yf[3, 4] = np.nan
    Yf[5, 6] = 0.5
    dev = GcateDevice(n, p, 3, 2, "nb", 100.0)
    assert dev.upload_counts(Yf) == (1, 1, 0)
    Y8 = np.minimum(Y, 255).astype(np.uint8)
    Y8[:, 7] = 0
    assert dev.upload_counts(Y8) == (0, 0, 1)
    dev.close()
