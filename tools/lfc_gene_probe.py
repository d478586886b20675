"""Follow-up of tools/config2_parity.py: the outcome model of LFC for a few named genes of the config-2
problem -- coefficients of the real device flow (`LFC`), of a direct device `fit_glm(Y, W, A)` (iteration
counts, guard flags) and of the oracle on the same design [X, U_device | A]."""
import json, os, sys, warnings
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    genes = [int(g) for g in sys.argv[1:]] or [1455, 2721, 0, 1, 2, 3]
    warnings.simplefilter("ignore")
    import causarray_b200 as cb
    from causarray_b200 import glm as cglm
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    from oracle import glm as og
    sim = simulate_screen(n_ctrl=800, n_perts=30, cells_per_pert=240, p=4000, r=5, d_cov=1, family="nb", seed=2,
                          base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    es = dict(rel_tol=2e-4, max_iters=30)
    r1, r2 = cb.fit_gcate(Y, X, A, 5, family="nb", disp_glm=sim["disp"], offset=True, kwargs_es_1=dict(es), kwargs_es_2=dict(es))
    ws = r2.pop("_ws")
    off = np.log(r2["kwargs_glm"]["size_factor"])
    W = np.c_[X, r2["U"]]
    df, est = cb.LFC(Y, W, pd.DataFrame(A), np.c_[X_A, r2["U"]], family="nb", offset=off, usevar="unequal",
                     _glm=ws.glm, _trusted=True)
    yh = est["Y_hat"]
    B_flow = np.c_[np.asarray(yh.Bcov), np.asarray(yh.Btrt)][genes]
    disp_g = np.asarray(cglm.estimate_disp(Y, np.c_[W, A], offset=off, disp_family="poisson", maxiter=1000, _glm=ws.glm))
    Bd, fitd, _, _, _ = cglm.fit_glm(Y, W, A, offset=off, family="nb", disp_glm=disp_g, maxiter=1000, _glm=ws.glm, _lazy=True)
    res = og.fit_glm_original(Y[:, genes], W, A, family="nb", disp_glm=disp_g[genes], offset=off, maxiter=1000)
    Bc, stc, itc = res[0], res[5], res[6]
    out = {"genes": genes,
           "iters_device_direct": np.asarray(fitd.n_iter)[genes].tolist(), "iters_oracle": np.asarray(itc).tolist(),
           "guard_device_direct": np.asarray(fitd.refit_mask)[genes].tolist(), "guard_oracle": (np.asarray(stc) != 0).tolist(),
           "B_flow_vs_oracle_abs_max": np.abs(B_flow - Bc).max(axis=1).tolist(),
           "B_direct_vs_oracle_abs_max": np.abs(Bd[genes] - Bc).max(axis=1).tolist(),
           "B_flow_vs_direct_abs_max": np.abs(B_flow - Bd[genes]).max(axis=1).tolist(),
           "max_abs_beta_trt_oracle": np.abs(Bc[:, W.shape[1]:]).max(axis=1).tolist(),
           "max_abs_beta_cov_oracle": np.abs(Bc[:, :W.shape[1]]).max(axis=1).tolist(),
           "argmax_B_flow_vs_oracle": np.abs(B_flow - Bc).argmax(axis=1).tolist()}
    ws.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
