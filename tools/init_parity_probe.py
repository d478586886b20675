"""Stage-by-stage comparison of the GCATE initialisation at config-2 scale (n = 8000), device path vs CPU
oracle on the same inputs: size factors, first per-gene GLM (coefficients, iteration counts, deviance
residuals), top-r SVD of the residuals (singular values, subspace distance; also LAPACK on the device
residuals to separate the SVD method from its input), second GLM on a common design.  One JSON line."""
import argparse, json, os, sys, time, warnings
import numpy as np
import scipy.linalg
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=1000)
    ap.add_argument("--perts", type=int, default=30)
    ap.add_argument("--cells", type=int, default=240)
    ap.add_argument("--ctrl", type=int, default=800)
    ap.add_argument("--r", type=int, default=5)
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    from causarray_b200 import glm as cglm
    from causarray_b200.optim import BatchWorkspace
    from causarray_b200.svd import topr_svd_device
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    from oracle import pipeline as op
    sim = simulate_screen(n_ctrl=args.ctrl, n_perts=args.perts, cells_per_pert=args.cells, p=args.p, r=args.r, d_cov=1,
                          family="nb", seed=2, base_lo=-1.5, base_hi=1.5)
    Y, A, disp = sim["Y"], sim["A"], sim["disp"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    Xf = np.c_[X, A]
    n, p = Y.shape
    d, r = Xf.shape[1], args.r
    cores = os.cpu_count() or 1
    out = {"n": n, "p": p, "d": d, "r": r}
    # --- device
    ws = BatchWorkspace.upload(Y, "nb", d, r)
    sf_g = ws.gcate.size_factors()
    ws.prepare(sf_g, disp)
    off_g = np.log(sf_g)
    B1g, fit, _, _, _ = cglm.fit_glm(Y, Xf, offset=off_g, family="nb", disp_glm=disp, maxiter=100, _glm=ws.glm, _lazy=True)
    it_g = np.asarray(fit.n_iter).copy()
    resid_g = np.array(ws.glm.resid(), copy=True)
    ug, sg, vtg = topr_svd_device(ws.glm.resid_device_ptr(), n, p, r, zero_cols=fit.refit_mask)
    # --- oracle
    sf_c = op.comp_size_factor(Y)
    off_c = np.log(sf_c)
    B1c, resid_c, st1, it_c, _ = op.fit_glm_parallel(Y, Xf, None, "nb", disp, off_c, maxiter=100, n_jobs=cores)
    uc, sc, vtc = op._svds_det(resid_c, r)
    out["size_factor_rel_max"] = float(np.max(np.abs(sf_g - sf_c) / sf_c))
    out["glm1_guard_genes"] = [int(np.sum(fit.refit_mask)), int(np.sum(st1 != 0))]
    ok = (~np.asarray(fit.refit_mask)) & (st1 == 0)
    out["glm1_B_abs_max"] = float(np.abs(B1g[ok] - B1c[ok]).max())
    out["glm1_B_abs_median_rowmax"] = float(np.median(np.abs(B1g[ok] - B1c[ok]).max(axis=1)))
    out["glm1_iter_equal_frac"] = float(np.mean(it_g[ok] == it_c[ok]))
    out["glm1_iter_diff_hist"] = {str(k): int(v) for k, v in zip(*np.unique((it_g[ok] - it_c[ok]), return_counts=True))}
    out["glm1_iter_median"] = [float(np.median(it_g[ok])), float(np.median(it_c[ok]))]
    same_it = ok & (it_g == it_c)
    if same_it.any():
        out["glm1_B_abs_max_same_iters"] = float(np.abs(B1g[same_it] - B1c[same_it]).max())
    out["resid_abs_max"] = float(np.abs(resid_g[:, ok] - resid_c[:, ok]).max())
    out["resid_abs_max_same_iters"] = float(np.abs(resid_g[:, same_it] - resid_c[:, same_it]).max()) if same_it.any() else None
    out["resid_fro_rel"] = float(np.linalg.norm(resid_g - resid_c) / np.linalg.norm(resid_c))
    out["svd_s_rel_max"] = float(np.max(np.abs(sg[::-1] - sc) / sc))       # device order is ascending
    proj = lambda U, V: float(np.linalg.norm(V - U @ (U.T @ V), 2))
    out["svd_subspace_dist_device_vs_oracle"] = proj(uc, ug)
    # LAPACK on the device residuals: is the distance due to the input or to the subspace iteration?
    Ul, sl, Vl = scipy.linalg.svd(np.where(np.asarray(fit.refit_mask)[None, :], 0.0, resid_g), full_matrices=False)
    out["svd_subspace_dist_device_vs_lapack_same_input"] = proj(Ul[:, :r], ug)
    out["svd_subspace_dist_lapack_deviceinput_vs_oracle"] = proj(uc, Ul[:, :r])
    out["svd_gap_ratio_s_r_plus_1_over_s_r"] = float(sl[r] / sl[r - 1])
    # second GLM on a COMMON design (the oracle's A0)
    Q1 = scipy.linalg.qr(Xf, mode="economic")[0]
    A0 = np.c_[Xf, uc - Q1 @ (Q1.T @ uc)]
    B0g, fit2, _, _, _ = cglm.fit_glm(Y, A0, offset=off_c, family="nb", disp_glm=disp, maxiter=100, _glm=ws.glm, _lazy=True)
    it2_g = np.asarray(fit2.n_iter).copy()
    B0c, _, st2, it2_c, _ = op.fit_glm_parallel(Y, A0, None, "nb", disp, off_c, maxiter=100, n_jobs=cores)
    ok2 = (~np.asarray(fit2.refit_mask)) & (st2 == 0)
    out["glm2_guard_genes"] = [int(np.sum(fit2.refit_mask)), int(np.sum(st2 != 0))]
    out["glm2_B_abs_max"] = float(np.abs(B0g[ok2] - B0c[ok2]).max())
    out["glm2_iter_equal_frac"] = float(np.mean(it2_g[ok2] == it2_c[ok2]))
    out["glm2_iter_diff_hist"] = {str(k): int(v) for k, v in zip(*np.unique((it2_g[ok2] - it2_c[ok2]), return_counts=True))}
    same2 = ok2 & (it2_g == it2_c)
    if same2.any():
        out["glm2_B_abs_max_same_iters"] = float(np.abs(B0g[same2] - B0c[same2]).max())
    if (~ok2).any():
        both = np.asarray(fit2.refit_mask) & (st2 != 0)
        if both.any():
            out["glm2_guard_B_abs_max"] = float(np.abs(B0g[both] - B0c[both]).max())
    ws.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
