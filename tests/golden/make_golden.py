"""Generate golden vectors from the REAL reference (build container only).

Run:  python tests/golden/make_golden.py
Needs ``/root/reference`` (imported through ``oracle/ref_shim.py``); writes
small ``.npz`` fixtures next to this file.  The fixtures travel to the GPU box,
the reference does not.

Every fixture records inputs AND the reference's outputs so both the CPU
oracle (``-m "not gpu"``) and the CUDA path (``-m gpu``) are checked against
the same numbers.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from causarray_b200.synth import simulate_screen  # noqa: E402


def _init_AB(rng, X, r, p, scale=0.1):
    n, d = X.shape
    A0 = np.c_[X, rng.standard_normal((n, r)) * scale]
    B0 = rng.standard_normal((p, d + r)) * scale
    return A0, B0


def golden_kernels(ca, family, seed, tag, C=None):
    """nll / grad / update_with_mask (stage 1 and stage 2) on one tiny problem."""
    import scipy.linalg
    from causarray.gcate_opt import update_with_mask, update
    from causarray.gcate_likelihood import nll, grad, gammaln_nb
    rng = np.random.default_rng(seed)
    sim = simulate_screen(n_ctrl=30, n_perts=3, cells_per_pert=10, p=70, r=2, d_cov=2, family=family, seed=seed)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    X = np.c_[Xc, Atr]
    n, p = Y.shape
    d, a, r = X.shape[1], Atr.shape[1], 2
    sf = np.exp(rng.standard_normal(n) * 0.2)
    nu = sim["disp"].copy()
    nu[::5] = 500.0                     # > thres_disp -> Poisson branch inside the NB family
    nu = nu.reshape(1, -1)
    thres = 100.0
    Yn = Y / sf[:, None]
    Ys = -gammaln_nb(Yn + 1)
    if family == "nb":
        Ys = Ys + np.where(nu > thres, 0.0, gammaln_nb(nu + Yn) - gammaln_nb(nu))
    A0, B0 = _init_AB(rng, X, r, p, scale=0.3)
    out = dict(Y=Y, X=X, sf=sf, nu=nu[0], A0=A0, B0=B0, d=d, a=a, r=r, thres_disp=thres, Ys=Ys)
    out["nll"] = nll(Yn, A0, B0, family, nu, Ys, thres)
    out["grad_B"] = grad(Yn, A0, B0, family, nu, thres)
    out["grad_A"] = grad(np.ascontiguousarray(Yn.T), B0, A0, family, np.ascontiguousarray(nu.T), thres)
    if C is None:
        C = 1e3 if family == "nb" else 1e5
    else:
        # a radius small enough that project_norm_ball (causarray/gcate_opt.py:10-26) rescales rows
        gA_, gB_ = out["grad_A"][:, d:], out["grad_B"][:, d:]
        print("  rows clipped at 2C = %g: cells %d / %d, genes %d / %d"
              % (2 * C, int((np.abs(gA_).max(1) > 2 * C).sum()), n, int((np.abs(gB_).max(1) > 2 * C).sum()), p))
    out["C"] = C
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    gene_active = rng.random(p) < 0.9
    cell_active = rng.random(n) < 0.9
    alpha_gene = np.where(rng.random(p) < 0.5, 0.1, 0.26)
    lam1 = np.zeros((p, d))
    # stage 1: three consecutive iterations
    A, B = A0.copy(), B0.copy()
    fs = []
    for it in range(3):
        f, A, B = update_with_mask(Yn, A, B, d, lam1, Q1, None, family, nu, Ys, thres, d, C,
                                   0.1, 0.5, 20, 1e-4, gene_active, cell_active, alpha_gene)
        fs.append(f)
        out["s1_A_%d" % it] = A.copy()
        out["s1_B_%d" % it] = B.copy()
    out["s1_f"] = np.array(fs)
    out["gene_active"], out["cell_active"], out["alpha_gene"], out["Q1"] = gene_active, cell_active, alpha_gene, Q1
    # stage 2 from the stage-1 result
    Q2 = scipy.linalg.qr(B[:, -r:], mode="economic")[0]
    B2 = B.copy()
    B2[:, d - a:d] -= Q2 @ (Q2.T @ B2[:, d - a:d])
    lam2 = 0.05 / (np.abs(B2[:, d - a:d]) + 1e-6)
    A2 = A.copy()
    out["s2_A_in"], out["s2_B_in"], out["lam2"], out["Q2"] = A2.copy(), B2.copy(), lam2, Q2
    fs = []
    for it in range(3):
        f, A2, B2 = update_with_mask(Yn, A2, B2, d, lam2, None, Q2, family, nu, Ys, thres, a, C,
                                     0.1, 0.5, 20, 1e-4, gene_active, cell_active, alpha_gene)
        fs.append(f)
        out["s2_A_%d" % it] = A2.copy()
        out["s2_B_%d" % it] = B2.copy()
    out["s2_f"] = np.array(fs)
    # neither P1 nor P2 (the T4 unit-test configuration), one iteration, all active
    A3, B3 = A0.copy(), B0.copy()
    f, A3, B3 = update(Yn, A3, B3, d, lam1, None, None, family, nu, Ys, thres, d, C, 0.1, 0.5, 20, 1e-4)
    out["s0_f"], out["s0_A"], out["s0_B"] = f, A3, B3
    np.savez_compressed(os.path.join(HERE, "kernels_%s.npz" % tag), **out)
    print("kernels", tag, "nll", out["nll"], "s1_f", out["s1_f"], "s2_f", out["s2_f"])


def golden_alter_min(ca, family, seed, tag):
    """Whole two-stage optimiser with explicit A0/B0 (skips GLM/SVD init)."""
    import scipy.linalg
    from causarray.gcate_opt import alter_min
    rng = np.random.default_rng(seed)
    sim = simulate_screen(n_ctrl=80, n_perts=3, cells_per_pert=25, p=180, r=3, d_cov=1, family=family, seed=seed)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    X = np.c_[Xc, Atr]
    n, p = Y.shape
    r, a = 3, Atr.shape[1]
    sf = ca.comp_size_factor(Y)
    disp = sim["disp"]
    A0, B0 = _init_AB(rng, X, r, p, scale=0.1)
    B0[:, 0] = np.log(Y.mean(0) + 0.1)
    kw = dict(family=family, disp_glm=disp.copy(), size_factor=sf.copy())
    res1 = alter_min(Y.copy(), r, X=X, P1=True, A=A0.copy(), B=B0.copy(), kwargs_glm=dict(kw),
                     kwargs_es=dict(max_iters=30))
    Q = scipy.linalg.qr(res1["B_Gamma"][:, -r:], mode="economic")[0]
    res2 = alter_min(Y.copy(), r, X=X, P2=Q, A=res1["X_U"].copy(), B=res1["B_Gamma"].copy(), lam=0.05, a=a,
                     kwargs_glm=dict(kw), kwargs_es=dict(max_iters=30))
    out = dict(Y=Y, X=X, a=a, r=r, sf=sf, disp=disp, A0=A0, B0=B0,
               s1_A=res1["X_U"], s1_B=res1["B_Gamma"], s1_hist=np.array(res1["hist"]), s1_n_iter=res1["n_iter"],
               s1_gene_updates=res1["total_gene_updates"], s1_cell_updates=res1["total_cell_updates"],
               s2_A=res2["X_U"], s2_B=res2["B_Gamma"], s2_hist=np.array(res2["hist"]), s2_n_iter=res2["n_iter"],
               s2_gene_updates=res2["total_gene_updates"], s2_cell_updates=res2["total_cell_updates"], Q2=Q)
    np.savez_compressed(os.path.join(HERE, "alter_min_%s.npz" % tag), **out)
    print("alter_min", tag, "iters", res1["n_iter"], res2["n_iter"], "hist1", res1["hist"][0], res1["hist"][-1],
          "hist2", res2["hist"][-1])


def golden_lfc(ca, seed, tag, usevar):
    """AIPW + LFC estimand + BH with Y_hat / pi_hat supplied (no GLM, no sklearn)."""
    rng = np.random.default_rng(seed)
    sim = simulate_screen(n_ctrl=60, n_perts=3, cells_per_pert=20, p=90, r=2, d_cov=1, family="nb", seed=seed)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    n, p = Y.shape
    a = Atr.shape[1]
    W = np.c_[Xc, sim["U"]]
    Bcov = np.c_[np.log(Y.mean(0) + 0.05), rng.standard_normal((p, 2)) * 0.1]
    Btrt = rng.standard_normal((p, a)) * 0.3
    sf = ca.comp_size_factor(Y)
    off = np.log(sf)
    Y[:, 0] = 0.0            # all-zero gene -> filtered
    Y[:, 1] = (rng.random(n) < 0.02) * 1.0   # very low expression
    eta0 = W @ Bcov.T + off[:, None]
    Y_hat = np.zeros((n, p, a, 2))
    for k in range(a):
        Y_hat[:, :, k, 0] = np.exp(eta0)
        Y_hat[:, :, k, 1] = np.exp(eta0 + Btrt[:, k][None, :])
    pi = np.clip(rng.uniform(0.05, 0.6, (n, a)), 0.01, 0.99)
    df, est = ca.LFC(Y, W, Atr, W_A=W, family="nb", offset=off, Y_hat=Y_hat.copy(), pi_hat=pi.copy(), usevar=usevar)
    from causarray.DR_estimation import AIPW_mean
    tau_aipw, _ = AIPW_mean(Y, np.stack([1 - Atr, Atr], -1), Y_hat, np.stack([1 - pi, pi], -1))
    cols = ["tau", "std", "log2fc", "log2fc_se", "stat", "rej", "pvalue", "padj", "pvalue_emp_null_adj",
            "padj_emp_null_adj", "mean_control", "mean_treated"]
    out = dict(Y=Y, W=W, A=Atr, Bcov=Bcov, Btrt=Btrt, offset=off, pi=pi, aipw_tau=tau_aipw,
               estimable=df["estimable"].to_numpy().astype(np.uint8), trt=df["trt"].to_numpy(),
               columns=np.array(list(df.columns)))
    for c in cols:
        out["df_" + c] = df[c].to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "lfc_%s.npz" % tag), **out)
    print("lfc", tag, "n sig", int(np.nansum(df["padj"] < 0.05)), "n filtered", int(np.isinf(df["std"]).sum()))


def golden_misc(ca):
    from causarray.utils import subsample_ctrl_cells, subsample_pert_cells
    from causarray.DR_inference import bh_correction
    from causarray.nb_glm_fast import _compute_nb_deviance_residuals, _compute_poisson_deviance_residuals
    from causarray.gcate_glm import estimate_disp
    rng = np.random.default_rng(7)
    sim = simulate_screen(n_ctrl=50, n_perts=2, cells_per_pert=20, p=60, r=2, seed=7)
    Y = sim["Y"]
    out = dict(Y=Y, size_factor=ca.comp_size_factor(Y))
    out["ctrl_sel"] = subsample_ctrl_cells(np.arange(5, 500, 3), n_ctrl=40, random_state=3)
    out["pert_sel"] = subsample_pert_cells(np.arange(2, 700, 5), max_cells=33, random_state=4)
    t = rng.standard_normal(200) * 1.7 + 0.2
    t[::17] = np.nan
    df = rng.uniform(3, 80, 200)
    out["bh_t"], out["bh_df"] = t, df
    for name, arr in zip(["bh_p", "bh_q", "bh_pa", "bh_qa"], bh_correction(t.copy(), df=df)):
        out[name] = arr
    for name, arr in zip(["bhz_p", "bhz_q", "bhz_pa", "bhz_qa"], bh_correction(t.copy(), df=None)):
        out[name] = arr
    mu = np.exp(rng.standard_normal(Y.shape) * 0.5) * (Y.mean(0) + 0.1)
    alpha = rng.uniform(0.2, 2.0, Y.shape[1])
    out["mu"], out["alpha"] = mu, alpha
    out["resid_nb"] = _compute_nb_deviance_residuals(Y, mu, alpha)
    out["resid_pois"] = _compute_poisson_deviance_residuals(Y, mu)
    off = np.log(out["size_factor"])
    out["disp_mom"] = estimate_disp(Y, Y_hat=(mu / np.exp(off)[:, None]).copy(), offset=off)
    Yc, A1, X1, XA1 = ca.prep_causarray_data(Y, sim["A"], X=None)
    out["prep_Y"], out["prep_X"], out["prep_XA"] = Yc, X1, XA1
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **out)
    print("misc ok")


def golden_pipeline(ca, family, seed, tag):
    """End-to-end fit_gcate + LFC through the REAL reference orchestration, with the
    statsmodels calls served by oracle/glm.py (parity unpinned for that piece)."""
    import causarray.gcate_glm as gg
    sim = simulate_screen(n_ctrl=150, n_perts=2, cells_per_pert=50, p=120, r=2, d_cov=1, family=family, seed=seed,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    disp = sim["disp"]
    # deterministic svds: fixed start vector, signs fixed by largest-abs entry
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    orig = sla.svds

    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    go.svds = svds_det
    go.sp.sparse.linalg.svds = svds_det
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    try:
        with gg._backend_override("original"):
            res1, res2 = ca.fit_gcate(Y.copy(), X, A, 2, family=family, disp_glm=disp.copy(),
                                      kwargs_es_1=dict(max_iters=30), kwargs_es_2=dict(max_iters=30))
            W = np.c_[X, res2["U"]]
            off = np.log(res2["kwargs_glm"]["size_factor"])
            df, est = ca.LFC(Y.copy(), W, A, W_A=W, family=family, offset=off, disp_glm=disp.copy(), backend="original")
    finally:
        go.svds = orig
        go.sp.sparse.linalg.svds = orig
        gg.Parallel = n_jobs_patch
    out = dict(Y=Y, X=X, A=A, disp=disp, s1_A=res1["X_U"], s1_B=res1["B_Gamma"], s1_hist=np.array(res1["hist"]),
               s2_A=res2["X_U"], s2_B=res2["B_Gamma"], s2_hist=np.array(res2["hist"]),
               size_factor=res2["kwargs_glm"]["size_factor"], pi_hat=est["pi_hat"], pi_hat_raw=est["pi_hat_raw"],
               Y_hat=est["Y_hat"].astype(np.float64), trt=df["trt"].to_numpy(),
               estimable=df["estimable"].to_numpy().astype(np.uint8))
    for c in ["tau", "std", "stat", "pvalue", "padj", "pvalue_emp_null_adj", "padj_emp_null_adj", "mean_control",
              "mean_treated", "log2fc", "log2fc_se"]:
        out["df_" + c] = df[c].to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "pipeline_%s.npz" % tag), **out)
    print("pipeline", tag, "iters", res1["n_iter"], res2["n_iter"], "sig", int(np.nansum(df["padj"] < 0.05)))


def golden_estimate_r(ca, seed):
    """JIC sweep (causarray/gcate.py:196-325) through the REAL reference orchestration."""
    import causarray.gcate_glm as gg
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    import causarray.gcate as gc
    sim = simulate_screen(n_ctrl=140, n_perts=2, cells_per_pert=40, p=110, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    disp = sim["disp"]
    orig = sla.svds

    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    go.svds = svds_det
    go.sp.sparse.linalg.svds = svds_det
    if hasattr(gc, "svds"):
        gc.svds = svds_det
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    try:
        with gg._backend_override("original"):
            df = ca.estimate_r(Y.copy(), X, A, 3, family="nb", disp_glm=disp.copy(),
                               kwargs_es_1=dict(max_iters=30), kwargs_es_2=dict(max_iters=30))
    finally:
        go.svds = orig
        go.sp.sparse.linalg.svds = orig
        if hasattr(gc, "svds"):
            gc.svds = orig
        gg.Parallel = n_jobs_patch
    out = dict(Y=Y, X=X, A=A, disp=disp, r=df["r"].to_numpy().astype(np.float64),
               deviance=df["deviance"].to_numpy().astype(np.float64), nu=df["nu"].to_numpy().astype(np.float64),
               JIC=df["JIC"].to_numpy().astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "estimate_r.npz"), **out)
    print("estimate_r", df.to_string())


def golden_lfc_fit(ca, seed, tag, cross_est):
    """LFC with its own nuisance fits (outcome GLM + logistic propensities) through the REAL
    reference, in-sample (K = 1) or cross-fitted (K = 2)."""
    import causarray.gcate_glm as gg
    sim = simulate_screen(n_ctrl=120, n_perts=2, cells_per_pert=40, p=70, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.5, base_hi=2.5)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    W = np.c_[Xc, sim["U"]]
    sf = ca.comp_size_factor(Y)
    off = np.log(sf)
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    try:
        with gg._backend_override("original"):
            df, est = ca.LFC(Y.copy(), W, Atr, W_A=W, family="nb", offset=off, disp_glm=sim["disp"].copy(),
                             cross_est=cross_est, usevar="unequal", backend="original")
    finally:
        gg.Parallel = n_jobs_patch
    out = dict(Y=Y, W=W, A=Atr, offset=off, disp=sim["disp"], pi_hat=est["pi_hat"], columns=np.array(list(df.columns)),
               estimable=df["estimable"].to_numpy().astype(np.uint8))
    for c in ["tau", "std", "stat", "pvalue", "padj", "mean_control", "mean_treated"]:
        out["df_" + c] = df[c].to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "lfc_fit_%s.npz" % tag), **out)
    print("lfc_fit", tag, "sig", int(np.nansum(df["padj"] < 0.05)), "filtered", int(np.isinf(df["std"]).sum()))


def golden_lfc_crossfit(ca, seed, tag, K, n_ctrl):
    """Proper cross-fitting (the reference's batched-backend semantics, ``DR_estimation.py:590-613``
    with ``nb_glm_fast.py:273-297``: fit on the training fold, predict the HELD-OUT rows with their
    own offsets) for fold shapes the statsmodels route cannot run (odd n, K = 3).  The fold-wise
    outcome model is the IRLS restatement (``oracle/glm.py``); everything downstream of ``Y_hat`` --
    cross-fitted propensities, AIPW, the estimand, BH -- is the REAL reference's ``LFC``."""
    from sklearn.model_selection import KFold
    from oracle import glm as og
    sim = simulate_screen(n_ctrl=n_ctrl, n_perts=2, cells_per_pert=40, p=60, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.5, base_hi=2.5)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    W = np.c_[Xc, sim["U"]]
    off = np.log(ca.comp_size_factor(Y))
    n, p = Y.shape
    a, d = Atr.shape[1], W.shape[1]
    Y_hat = np.zeros((n, p, a, 2))
    for tr, te in KFold(n_splits=K, random_state=0, shuffle=True).split(W):
        B = og.fit_glm_original(Y[tr], W[tr], Atr[tr], family="nb", disp_glm=sim["disp"], offset=off[tr], maxiter=1000)[0]
        eta0 = W[te] @ B[:, :d].T + off[te][:, None]
        Y_hat[te, :, :, 0] = np.exp(eta0)[:, :, None]
        Y_hat[te, :, :, 1] = np.exp(eta0[:, :, None] + B[:, d:][None, :, :])
    df, est = ca.LFC(Y.copy(), W, Atr, W_A=W, family="nb", offset=off, Y_hat=Y_hat.copy(), K=K, usevar="unequal")
    out = dict(Y=Y, W=W, A=Atr, offset=off, disp=sim["disp"], K=K, pi_hat=est["pi_hat"], Y_hat=np.clip(Y_hat, None, 1e5),
               estimable=df["estimable"].to_numpy().astype(np.uint8))
    for c in ["tau", "std", "stat", "pvalue", "padj", "mean_control", "mean_treated"]:
        out["df_" + c] = df[c].to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "lfc_crossfit_%s.npz" % tag), **out)
    print("lfc_crossfit", tag, "n", n, "K", K, "sig", int(np.nansum(df["padj"] < 0.05)))


def golden_fdx(ca, seed):
    """FDX control (multiplier bootstrap + step-down + augmentation, causarray/DR_inference.py:8-150)
    from the REAL reference: the function on synthetic statistics (ties, untestable hypotheses, several
    alpha / c) and LFC(fdx=True) end to end with the nuisances supplied.  The bootstrap draws come from
    the global numpy stream, seeded right before each call."""
    from causarray import DR_inference as ri
    rng = np.random.default_rng(seed)
    out = {}
    cases = []
    for ci, (n, p, nsig, alpha, c, B) in enumerate([(150, 200, 25, 0.05, 0.1, 300), (150, 200, 60, 0.1, 0.3, 300),
                                                    (80, 120, 0, 0.05, 0.1, 200), (150, 200, 40, 0.05, 0.0, 300)]):
        std = np.abs(rng.standard_normal(p)) * 0.1 + 0.05
        eta = rng.standard_normal((n, p)) * std[None, :] / np.sqrt(n)
        tau = rng.standard_normal(p) * std
        idx = rng.choice(p, nsig, replace=False)
        tau[idx] += rng.choice([-1.0, 1.0], nsig) * rng.uniform(4, 9, nsig) * std[idx]
        if ci == 1:
            std[3:7] = 1e-9                        # below min_var: not tested, discovery iff |tau| > min_diff
            tau[3], tau[4] = 0.5, 0.01
            tau[10] = tau[11] = 6 * std[10]
            std[11] = std[10]                      # an exact tie in |t|
        t = tau / std
        np.random.seed(100 + ci)
        V = ri.fdx_control(tau, t.copy(), eta, std, True, B, alpha, c)
        out.update({"c%d_tau" % ci: tau, "c%d_t" % ci: t, "c%d_eta" % ci: eta, "c%d_std" % ci: std,
                    "c%d_par" % ci: np.array([B, alpha, c, 100 + ci]), "c%d_V" % ci: V})
        cases.append(int(V.sum()))
    sim = simulate_screen(n_ctrl=200, n_perts=2, cells_per_pert=100, p=80, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=1.0, base_hi=3.0, frac_nonnull=0.25)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    W = np.c_[Xc, sim["U"]]
    sf = ca.comp_size_factor(Y)
    n, p = Y.shape
    a = Atr.shape[1]
    # outcome model close to the truth: control-cell mean per gene, true effects, size factors
    ctrl = Atr.sum(1) == 0
    mu0 = (Y[ctrl] / sf[ctrl, None]).mean(0)[None, :] * sf[:, None] + 1e-3
    Y_hat = np.stack([np.repeat(mu0[:, :, None], a, axis=2), mu0[:, :, None] * np.exp(sim["tau"].T[None, :, :])], axis=-1)
    pi = np.clip(np.repeat(Atr.mean(0)[None, :] / (Atr.mean(0)[None, :] + ctrl.mean()), n, axis=0)
                 * rng.uniform(0.8, 1.2, (n, a)), 0.01, 0.99)
    np.random.seed(321)
    df, est = ca.LFC(Y.copy(), W, Atr, W_A=W, family="nb", offset=np.log(sf), Y_hat=Y_hat.copy(), pi_hat=pi.copy(),
                     usevar="unequal", fdx=True, fdx_alpha=0.1, fdx_c=0.2)
    out.update(dict(Y=Y, W=W, A=Atr, offset=np.log(sf), Y_hat=Y_hat, pi=pi, lfc_rej=df["rej"].to_numpy().astype(np.float64),
                    lfc_tau=df["tau"].to_numpy().astype(np.float64), lfc_padj=df["padj"].to_numpy().astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "fdx.npz"), **out)
    print("fdx discoveries per case", cases, "LFC rej", int(df["rej"].sum()))


def golden_more(ca, seed):
    """fit_gcate_batch with pert-cell subsampling and the warm start of U across batches, and
    fit_glm with its own dispersion estimate and imputation, through the REAL reference."""
    import causarray.gcate_glm as gg
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    sim = simulate_screen(n_ctrl=150, n_perts=4, cells_per_pert=40, p=90, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    orig = sla.svds

    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    go.svds = svds_det
    go.sp.sparse.linalg.svds = svds_det
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    out = dict(Y=Y, X=X, A=A, disp=sim["disp"])
    try:
        with gg._backend_override("original"):
            res = ca.fit_gcate_batch(Y.copy(), X, A, 2, batch_size=2, max_cells=60, n_ctrl=100, family="nb",
                                     disp_glm=sim["disp"].copy(), warm_start_U=True, random_state=3,
                                     kwargs_es_1=dict(max_iters=15), kwargs_es_2=dict(max_iters=15))
            out["n_batches"] = len(res)
            for i, b in enumerate(res):
                out["b%d_cell_idx" % i] = np.asarray(b["cell_idx"])
                out["b%d_ctrl_idx" % i] = np.asarray(b["ctrl_idx"])
                out["b%d_pert_idx" % i] = np.asarray(b["pert_idx"])
                out["b%d_X_U" % i] = b["res_2"]["X_U"]
                out["b%d_B_Gamma" % i] = b["res_2"]["B_Gamma"]
                out["b%d_hist1" % i] = np.array(b["res_1"]["hist"])
                out["b%d_hist2" % i] = np.array(b["res_2"]["hist"])
            B, Yhat, disp, offsets, resid = ca.fit_glm(Y.copy(), X, A, family="nb", disp_glm=None, offset=True,
                                                       impute=True, maxiter=100)
            out.update(glm_B=B, glm_Yhat0=Yhat[0], glm_Yhat1=Yhat[1], glm_disp=disp, glm_offsets=offsets,
                       glm_resid=resid)
    finally:
        go.svds = orig
        go.sp.sparse.linalg.svds = orig
        gg.Parallel = n_jobs_patch
    np.savez_compressed(os.path.join(HERE, "more.npz"), **out)
    print("more: batches", len(res), [len(b["cell_idx"]) for b in res], "glm B", B.shape)


def golden_lfc_options(ca, seed):
    """LFC option variants through the REAL reference: pooled variance with a cell mask and
    unclipped propensities (NB, dispersion estimated inside), and the Poisson family."""
    import causarray.gcate_glm as gg
    sim = simulate_screen(n_ctrl=130, n_perts=3, cells_per_pert=30, p=60, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.5, base_hi=2.5)
    Y, Xc, Atr = sim["Y"], sim["X"], sim["A"]
    rng = np.random.default_rng(seed)
    W = np.c_[Xc, sim["U"]]
    n = Y.shape[0]
    ctrl = Atr.sum(1) == 0
    mask = np.zeros(Atr.shape, dtype=bool)
    for j in range(Atr.shape[1]):
        mask[:, j] = (Atr[:, j] == 1) | (ctrl & (rng.random(n) < 0.7))
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    out = dict(Y=Y, W=W, A=Atr, mask=mask)
    try:
        with gg._backend_override("original"):
            df1, est1 = ca.LFC(Y.copy(), W, Atr, W_A=W, family="nb", offset=True, mask=mask.copy(), usevar="pooled",
                               ps_clip=None, thres_min=0.05, thres_diff=0.02, backend="original")
            df2, est2 = ca.LFC(Y.copy(), W, Atr, W_A=W, family="poisson", offset=True, usevar="unequal",
                               backend="original")
    finally:
        gg.Parallel = n_jobs_patch
    for tag, df, est in (("v1", df1, est1), ("v2", df2, est2)):
        out[tag + "_pi_hat"] = est["pi_hat"]
        out[tag + "_estimable"] = df["estimable"].to_numpy().astype(np.uint8)
        for c in ["tau", "std", "stat", "pvalue", "padj", "mean_control", "mean_treated"]:
            out[tag + "_" + c] = df[c].to_numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "lfc_options.npz"), **out)
    print("lfc_options sig", int(np.nansum(df1["padj"] < 0.05)), int(np.nansum(df2["padj"] < 0.05)),
          "filtered", int(np.isinf(df1["std"]).sum()), int(np.isinf(df2["std"]).sum()))


def golden_gcate_options(ca, seed):
    """fit_gcate with non-default options through the REAL reference: no offset, explicit c1,
    non-default line-search and early-stopping settings.  (The dispersion is passed in: with
    ``disp_glm=None`` and the statsmodels backend the reference's ``estimate_disp_auto`` falls off its
    end and every gene gets nu = 1 -- a documented deviation)"""
    import causarray.gcate_glm as gg
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    sim = simulate_screen(n_ctrl=120, n_perts=2, cells_per_pert=45, p=80, r=3, d_cov=2, family="nb", seed=seed,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    orig = sla.svds

    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    go.svds = svds_det
    go.sp.sparse.linalg.svds = svds_det
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    kw = dict(family="nb", disp_glm=sim["disp"].copy(), offset=False, c1=0.2,
              kwargs_ls_1=dict(alpha=0.05, beta=0.6, tol=1e-3), kwargs_ls_2=dict(alpha=0.2, max_iters=15),
              kwargs_es_1=dict(max_iters=25, patience=3, rel_tol=1e-5), kwargs_es_2=dict(max_iters=12, tolerance=1e-6))
    try:
        with gg._backend_override("original"):
            res1, res2 = ca.fit_gcate(Y.copy(), X, A, 3, **kw)
    finally:
        go.svds = orig
        go.sp.sparse.linalg.svds = orig
        gg.Parallel = n_jobs_patch
    out = dict(Y=Y, X=X, A=A, s1_A=res1["X_U"], s1_B=res1["B_Gamma"], s1_hist=np.array(res1["hist"]),
               s2_A=res2["X_U"], s2_B=res2["B_Gamma"], s2_hist=np.array(res2["hist"]),
               disp=np.ravel(res2["kwargs_glm"]["disp_glm"]), size_factor=np.ravel(res2["kwargs_glm"]["size_factor"]))
    np.savez_compressed(os.path.join(HERE, "gcate_options.npz"), **out)
    print("gcate_options iters", res1["n_iter"], res2["n_iter"], len(res1["hist"]), len(res2["hist"]))


def golden_batch(ca, seed):
    """gcate_lfc_batch (causarray/DR_learner.py:584-866) through the REAL reference: control
    subsampling, per-batch fits (dispersion estimated once, up front), result frame."""
    import causarray.gcate_glm as gg
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    sim = simulate_screen(n_ctrl=160, n_perts=4, cells_per_pert=35, p=100, r=2, d_cov=1, family="nb", seed=seed,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    orig = sla.svds

    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    go.svds = svds_det
    go.sp.sparse.linalg.svds = svds_det
    n_jobs_patch = gg.Parallel
    gg.Parallel = lambda n_jobs=None: (lambda gen: [f(*a, **k) for f, a, k in gen])
    kw = dict(batch_size=2, max_cells=2000, n_ctrl=120, family="nb",
              gcate_kwargs=dict(disp_glm=sim["disp"].copy(), kwargs_es_1=dict(max_iters=20), kwargs_es_2=dict(max_iters=20)),
              lfc_kwargs=dict(usevar="unequal", disp_glm=sim["disp"].copy()), random_state=0)
    try:
        with gg._backend_override("original"):
            df = ca.gcate_lfc_batch(Y.copy(), X, A, 2, **kw)
    finally:
        go.svds = orig
        go.sp.sparse.linalg.svds = orig
        gg.Parallel = n_jobs_patch
    out = dict(Y=Y, X=X, A=A, disp=sim["disp"], columns=np.array(list(df.columns)))
    for c in df.columns:
        v = df[c].to_numpy()
        out["df_" + c] = v.astype(str) if v.dtype == object else v.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "batch.npz"), **out)
    print("batch", df.shape, list(df.columns), "sig", int(np.nansum(df["padj"] < 0.05)))


if __name__ == "__main__":
    ca = ref_shim.import_reference()
    which = sys.argv[1:] or ["kernels", "alter_min", "lfc", "misc", "pipeline", "estimate_r", "batch", "lfc_fit", "more", "lfc_options", "gcate_options", "lfc_crossfit", "fdx", "kernels_clip"]
    if "estimate_r" in which:
        golden_estimate_r(ca, 41)
    if "batch" in which:
        golden_batch(ca, 51)
    if "gcate_options" in which:
        golden_gcate_options(ca, 91)
    if "lfc_options" in which:
        golden_lfc_options(ca, 81)
    if "more" in which:
        golden_more(ca, 71)
    if "fdx" in which:
        golden_fdx(ca, 95)
    if "lfc_crossfit" in which:
        golden_lfc_crossfit(ca, 63, "k2_odd", 2, 121)
        golden_lfc_crossfit(ca, 64, "k3", 3, 120)
    if "lfc_fit" in which:
        golden_lfc_fit(ca, 61, "k1", False)
        golden_lfc_fit(ca, 62, "k2", True)
    if "kernels" in which:
        golden_kernels(ca, "nb", 11, "nb")
        golden_kernels(ca, "poisson", 12, "poisson")
    if "kernels_clip" in which:
        golden_kernels(ca, "nb", 13, "nb_clip", C=0.01)
    if "alter_min" in which:
        golden_alter_min(ca, "nb", 21, "nb")
        golden_alter_min(ca, "poisson", 22, "poisson")
    if "lfc" in which:
        golden_lfc(ca, 31, "unequal", "unequal")
        golden_lfc(ca, 32, "pooled", "pooled")
    if "misc" in which:
        golden_misc(ca)
    if "pipeline" in which:
        golden_pipeline(ca, "nb", 41, "nb")
        golden_pipeline(ca, "poisson", 42, "poisson")
