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
