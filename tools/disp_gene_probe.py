"""Follow-up of tools/config2_parity.py: the Poisson fit + moments dispersion behind LFC for a few named
genes of the config-2 problem, device vs oracle, on the design [X, U, A] of the device fit."""
import json, os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    genes = [int(g) for g in sys.argv[1:]] or [1455, 2721, 0, 1, 2, 3]
    warnings.simplefilter("ignore")
    import causarray_b200 as cb
    from causarray_b200 import glm as cglm
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    from oracle import pipeline as op, glm as og
    sim = simulate_screen(n_ctrl=800, n_perts=30, cells_per_pert=240, p=4000, r=5, d_cov=1, family="nb", seed=2,
                          base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    es = dict(rel_tol=2e-4, max_iters=30)
    r1, r2 = cb.fit_gcate(Y, X, A, 5, family="nb", disp_glm=sim["disp"], offset=True, kwargs_es_1=dict(es), kwargs_es_2=dict(es))
    ws = r2.pop("_ws")
    off = np.log(r2["kwargs_glm"]["size_factor"])
    Xf = np.c_[X, r2["U"], A]
    Bp, fitp, _, _, _ = cglm.fit_glm(Y, Xf, offset=off, family="poisson", maxiter=1000, _glm=ws.glm, _lazy=True)
    it_g, refit_g = np.asarray(fitp.n_iter)[genes].tolist(), np.asarray(fitp.refit_mask)[genes].tolist()
    disp_g = np.asarray(cglm.estimate_disp(Y, Xf, offset=off, disp_family="poisson", maxiter=1000, _glm=ws.glm))[genes]
    Bc, _, stc, itc, mu = op.fit_glm_parallel(Y[:, genes], Xf, None, "poisson", None, off, maxiter=1000, n_jobs=1)
    disp_c = og.estimate_disp_moments(Y[:, genes], mu, off)
    Yn = Y[:, genes] / np.exp(off)[:, None]
    Yh = mu / np.exp(off)[:, None]
    out = {"genes": genes, "poisson_iters_device": it_g, "poisson_iters_oracle": np.asarray(itc).tolist(),
           "poisson_guard_device": refit_g, "poisson_guard_oracle": (np.asarray(stc) != 0).tolist(),
           "poisson_B_abs_max": np.abs(Bp[genes] - Bc).max(axis=1).tolist(),
           "disp_device": disp_g.tolist(), "disp_oracle": disp_c.tolist(),
           "disp_rel": (np.abs(disp_g - disp_c) / disp_c).tolist(),
           "phi_oracle": (np.mean((Yn - np.minimum(Yh, Yn.max(0))) ** 2 - np.minimum(Yh, Yn.max(0)), axis=0)
                          / np.mean(np.minimum(Yh, Yn.max(0)) ** 2, axis=0)).tolist(),
           "fitted_exceeds_max_count_cells": (Yh > Yn.max(0)[None, :]).sum(axis=0).tolist(),
           "zero_fraction": (Y[:, genes] == 0).mean(axis=0).tolist()}
    ws.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
