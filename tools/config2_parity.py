"""BASELINE config 2 (n = 8000 cells: 800 controls + 30 perturbations x 240, p = 4000 genes, r = 5, NB;
k = 36) through the public API on one GPU, against the CPU oracle (oracle/pipeline.py: C/OpenMP alternating
minimisation + per-gene numpy IRLS over the box's cores) on the SAME inputs: `fit_gcate` + `LFC`, no
batching.  Prints one JSON line: relative errors of tau / std, agreement of the BH rejection sets at
alpha = 0.05, iteration counts and wall times of both sides.

    python tools/config2_parity.py [--p 4000] [--perts 30] [--cells 240] [--ctrl 800]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=4000)
    ap.add_argument("--perts", type=int, default=30)
    ap.add_argument("--cells", type=int, default=240)
    ap.add_argument("--ctrl", type=int, default=800)
    ap.add_argument("--r", type=int, default=5)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--label", default="C2")
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    from oracle import pipeline as op, cport
    sim = simulate_screen(n_ctrl=args.ctrl, n_perts=args.perts, cells_per_pert=args.cells, p=args.p, r=args.r, d_cov=1,
                          family="nb", seed=args.seed, base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    n, p = Y.shape
    es = dict(rel_tol=2e-4, max_iters=args.iters)
    t0 = time.perf_counter()
    r1, r2 = cb.fit_gcate(Y, X, A, args.r, family="nb", disp_glm=sim["disp"], offset=True, kwargs_es_1=dict(es),
                          kwargs_es_2=dict(es))
    ws = r2.pop("_ws")
    off = np.log(r2["kwargs_glm"]["size_factor"])
    df, _ = cb.LFC(Y, np.c_[X, r2["U"]], pd.DataFrame(A), np.c_[X_A, r2["U"]], family="nb", offset=off,
                   usevar="unequal", _glm=ws.glm, _trusted=True)
    t_gpu = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    cport.set_threads(cores)
    t0 = time.perf_counter()
    g, l, st = op.gcate_lfc_one_batch_cpu(Y, X, A, X_A, args.r, family="nb", disp=sim["disp"], es1=dict(es),
                                          es2=dict(es), usevar="unequal", n_jobs=cores)
    t_cpu = time.perf_counter() - t0
    a = A.shape[1]
    tau_g = df["tau"].to_numpy().reshape(a, p)
    std_g = df["std"].to_numpy().reshape(a, p)
    padj_g = df["padj"].to_numpy().reshape(a, p)
    tau_c, std_c, padj_c = l["tau"], np.sqrt(l["var"]), l["padj"]
    fin = np.isfinite(std_c) & np.isfinite(std_g) & (tau_c != 0)
    rel = lambda x, y: np.abs(x - y) / np.maximum(np.abs(y), 1e-12)
    rt, rs = rel(tau_g[fin], tau_c[fin]), rel(std_g[fin], std_c[fin])
    rej_g, rej_c = padj_g < 0.05, padj_c < 0.05
    out = {
        "config": args.label, "n": n, "p": p, "a": a, "r": args.r, "k": int(r2["X_U"].shape[1]),
        "gpu_seconds": round(t_gpu, 3), "cpu_seconds": round(t_cpu, 1), "cpu_cores": cores,
        "iters_gpu": [len(r1["hist"]) - 1, len(r2["hist"]) - 1], "iters_cpu": [len(g["hist1"]) - 1, len(g["hist2"]) - 1],
        "pairs_compared": int(fin.sum()), "pairs_total": int(a * p),
        "same_nonestimable_pattern": bool(np.array_equal(np.isfinite(std_g), np.isfinite(std_c))),
        "tau_rel_max": float(rt.max()), "tau_rel_p999": float(np.quantile(rt, 0.999)), "tau_rel_median": float(np.median(rt)),
        "std_rel_max": float(rs.max()), "std_rel_p999": float(np.quantile(rs, 0.999)), "std_rel_median": float(np.median(rs)),
        "tau_abs_max": float(np.abs(tau_g[fin] - tau_c[fin]).max()),
        "rejections_gpu": int(rej_g.sum()), "rejections_cpu": int(rej_c.sum()), "rejections_differ": int((rej_g != rej_c).sum()),
    }
    # where the differences sit: genes that took the regularised fallback in the outcome model, the
    # objective histories (initialisation = first entry), and the latent factors after matching columns
    guard = np.asarray(l["status"]) != 0
    gm = np.broadcast_to(guard[None, :], fin.shape)
    for name, sel in (("guard_genes", fin & gm), ("clean_genes", fin & ~gm)):
        if sel.any():
            rr = rel(tau_g[sel], tau_c[sel])
            out["tau_rel_" + name] = {"n": int(sel.sum()), "max": float(rr.max()), "p999": float(np.quantile(rr, 0.999)),
                                      "median": float(np.median(rr)),
                                      "abs_max": float(np.abs(tau_g[sel] - tau_c[sel]).max())}
    out["n_guard_genes_lfc"] = int(guard.sum())
    h1g, h1c, h2g, h2c = (np.asarray(x, dtype=float) for x in (r1["hist"], g["hist1"], r2["hist"], g["hist2"]))
    if len(h1g) == len(h1c) and len(h2g) == len(h2c):
        out["hist1_rel_first_last_max"] = [float(abs(h1g[0] - h1c[0]) / abs(h1c[0])), float(abs(h1g[-1] - h1c[-1]) / abs(h1c[-1])),
                                           float(np.max(np.abs(h1g - h1c) / np.abs(h1c)))]
        out["hist2_rel_first_last_max"] = [float(abs(h2g[0] - h2c[0]) / abs(h2c[0])), float(abs(h2g[-1] - h2c[-1]) / abs(h2c[-1])),
                                           float(np.max(np.abs(h2g - h2c) / np.abs(h2c)))]
    Ug, Uc = r2["U"], g["U"]
    C = Ug.T @ Uc
    match = np.argmax(np.abs(C), axis=1)
    sign = np.sign(C[np.arange(C.shape[0]), match])
    out["U_max_abs_diff_matched_columns"] = float(np.abs(Ug - Uc[:, match] * sign[None, :]).max())
    Bg_, Bc_ = r2["B_Gamma"], g["B_Gamma"]
    d_ = X.shape[1] + a
    out["B_trt_max_abs_diff"] = float(np.abs(Bg_[:, X.shape[1]:d_] - Bc_[:, X.shape[1]:d_]).max())
    # the pairs above 1e-8: which genes, and does the outcome GLM of those genes stop at the same
    # iteration on both sides?  (same design, offset and dispersion on both sides: the oracle's)
    from causarray_b200 import glm as cglm
    worst = np.zeros((a, p), dtype=bool)
    worst[fin] = rt > 1e-8
    wg = np.where(worst.any(axis=0))[0]
    out["genes_with_pairs_above_1e-8"] = [int(j) for j in wg[:20]]
    out["pairs_above_1e-8"] = int(worst.sum())
    Wc = np.c_[X, g["U"]]
    offc = np.log(g["size_factor"])
    Bg2, fit2, _, _, _ = cglm.fit_glm(Y, np.c_[Wc, A], offset=offc, family="nb", disp_glm=l["disp"], maxiter=1000,
                                      _glm=ws.glm, _lazy=True)
    itg = np.asarray(fit2.n_iter)
    Bc2, _, stc, itc, _ = op.fit_glm_parallel(Y, np.c_[Wc, A], None, "nb", l["disp"], offc, maxiter=1000, n_jobs=cores)
    okg = (~np.asarray(fit2.refit_mask)) & (stc == 0)
    diff_it = okg & (itg != itc)
    out["outcome_glm_iter_mismatch_genes"] = [int(j) for j in np.where(diff_it)[0][:20]]
    out["outcome_glm_iter_mismatch_count"] = int(diff_it.sum())
    out["outcome_glm_B_abs_max_same_iters"] = float(np.abs(Bg2[okg & ~diff_it] - Bc2[okg & ~diff_it]).max())
    if diff_it.any():
        out["outcome_glm_B_abs_max_mismatch"] = float(np.abs(Bg2[diff_it] - Bc2[diff_it]).max())
        out["outcome_glm_iters_mismatch"] = [[int(itg[j]), int(itc[j])] for j in np.where(diff_it)[0][:20]]
    out["worst_genes_are_iter_mismatch_genes"] = bool(set(wg.tolist()) <= set(np.where(diff_it)[0].tolist()))
    ws.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
