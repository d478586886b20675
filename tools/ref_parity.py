"""Full-size parity of the device path against the REAL reference (staged oracle/_ref: numba
alter_min, the reference's own fit_gcate + LFC; only the statsmodels calls are served by
oracle/shims) on the same inputs, with the CPU timing of the reference at the full gene count.

    python tools/ref_parity.py --config 3 [--seed 100] [--out gpurun_out/x.json]

config 1: simu_poi shape (1000 x 2000, Poisson, r = 2, one binary treatment, 3 covariates)
config 2: 8000 cells x 4000 genes, 30 perturbations, r = 5, NB (fit_gcate + LFC, no batching)
config 3: one batch of the benchmark workload (bench.make_batch: 3995 x 8000, 15 perturbations, r = 10)

Besides tau / std / rejection agreement the record lists, for every gene with a pair above 1e-8,
the IRLS iteration counts of ALL FOUR per-gene fits (both GCATE initialisation fits, the Poisson
pre-fit of the dispersion estimate, the NB outcome fit) on both sides and the last |dev change|
values around the stopping decision (statsmodels' rule: |dev_k - dev_{k-1}| <= 1e-8, absolute).
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_data(cfg, seed):
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    if cfg == 3:
        import bench
        b = bench.make_batch(seed)
        return dict(Y=b["Y"], X=b["X"], A=b["A"], X_A=b["X_A"], disp=b["disp"], r=10, family="nb", iters=30)
    if cfg == 2:
        sim = simulate_screen(n_ctrl=800, n_perts=30, cells_per_pert=240, p=4000, r=5, d_cov=1, family="nb", seed=seed,
                              base_lo=-1.5, base_hi=1.5)
        fam, r = "nb", 5
    elif cfg == 1:
        sim = simulate_screen(n_ctrl=500, n_perts=1, cells_per_pert=500, p=2000, r=2, d_cov=3, family="poisson", seed=seed,
                              base_lo=0.0, base_hi=2.5)
        fam, r = "poisson", 2
    else:
        raise SystemExit("config must be 1, 2 or 3")
    Y, A = sim["Y"], sim["A"]
    X = sim["X"]
    X_A = np.c_[X, (np.log(Y.sum(1)) - np.log(Y.sum(1)).mean()) / np.log(Y.sum(1)).std()]
    return dict(Y=Y, X=X, A=A, X_A=X_A, disp=sim["disp"], r=r, family=fam, iters=30)


def annotate(out, tol=1e-8):
    """Stop-rule summary of the probed genes.  A fit is BORDERLINE when a deviance change next to its stopping
    decision lies within 1e-11 of the threshold: the decision then hangs on the rounding of a sum of n unit
    deviances (~1e-13 .. 1e-12 absolute), on either side -- the reference's own arithmetic (its numpy sums, its
    least-squares solver) decides such a gene differently from a restatement of the same formulas, too."""
    flips, border = [], []
    for q in out.get("probes", []):
        if any(v["iters_device"] != v["iters_cpu"] for v in q["fits"].values()):
            flips.append(q["gene"])
        margins = {}
        for name, v in q["fits"].items():
            ch = [c for c in list(v["cpu_last_dev_changes"]) + list(v["device_dev_changes"]) if c > 0.0]
            if ch:
                margins[name] = min(abs(c - tol) for c in ch)
        q["stop_margin"] = margins
        if margins and min(margins.values()) < 1e-11:
            border.append({"gene": q["gene"], "fit": min(margins, key=margins.get), "margin": min(margins.values())})
    out["probed_genes_with_a_stop_rule_flip"] = flips
    out["probed_genes_with_a_borderline_stop"] = border
    out["probed_genes_without_flip"] = [q["gene"] for q in out.get("probes", []) if q["gene"] not in flips]
    d = out.get("irls_gene_iters_ref", 0) - out.get("irls_gene_iters_gpu", 0)
    out["irls_gene_iters_ref_minus_gpu"] = d
    return out


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--annotate":   # re-derive the stop-rule summary of a saved record
        with open(sys.argv[2]) as fh:
            rec = json.loads(fh.read())
        with open(sys.argv[2], "w") as fh:
            fh.write(json.dumps(annotate(rec)) + "\n")
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-probe", type=int, default=12)
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    import causarray_b200 as cb
    from causarray_b200 import glm as cglm
    from causarray_b200.device import GlmDevice
    from oracle import refdrive, glm as og

    D = make_data(args.config, args.seed)
    Y, X, A, X_A, disp, r, fam = D["Y"], D["X"], D["A"], D["X_A"], D["disp"], D["r"], D["family"]
    n, p = Y.shape
    a = A.shape[1]
    es = dict(rel_tol=2e-4, max_iters=D["iters"])

    # ---- device side; every per-gene fit is recorded (design, offset, dispersion, iteration counts) ----
    fits = []
    _fit = GlmDevice.fit

    def fit_rec(self, family, Xd, offset=None, nu=None, **k):
        out = _fit(self, family, Xd, offset, nu, **k)
        if k.get("gene_mask") is None:
            fits.append(dict(family=family, X=np.array(Xd), offset=None if offset is None else np.array(offset),
                             nu=None if nu is None else np.array(nu), maxiter=k.get("maxiter", 100),
                             n_iter=np.array(out[2]), status=np.array(out[3]), B=np.array(out[0])))
        return out
    GlmDevice.fit = fit_rec
    t0 = time.perf_counter()
    r1, r2 = cb.fit_gcate(Y, X, A, r, family=fam, disp_glm=disp, offset=True, kwargs_es_1=dict(es), kwargs_es_2=dict(es))
    ws = r2.pop("_ws")
    off = np.log(r2["kwargs_glm"]["size_factor"])
    df, _ = cb.LFC(Y, np.c_[X, r2["U"]], pd.DataFrame(A), np.c_[X_A, r2["U"]], family=fam, offset=off,
                   usevar="unequal", _glm=ws.glm, _trusted=True)
    t_gpu = time.perf_counter() - t0
    GlmDevice.fit = _fit
    fit_names = ["init1", "init2", "lfc_poisson_prefit", "lfc_outcome"] if fam == "nb" else ["init1", "init2", "lfc_outcome"]

    # ---- the reference ----
    jit = refdrive.warm_up_jit(fam)
    q1, q2, dfr, estr, st = refdrive.reference_fit_gcate_lfc(Y, X, A, X_A, r, family=fam, disp=disp, es1=dict(es),
                                                             es2=dict(es), usevar="unequal")
    tau_g, std_g, padj_g = (df[c].to_numpy().reshape(a, p) for c in ("tau", "std", "padj"))
    tau_c, std_c, padj_c = (dfr[c].to_numpy().reshape(a, p) for c in ("tau", "std", "padj"))
    fin = np.isfinite(std_c) & np.isfinite(std_g) & (tau_c != 0)
    rel = lambda x, y: np.abs(x - y) / np.maximum(np.abs(y), 1e-12)
    rt = np.zeros((a, p))
    rs = np.zeros((a, p))
    rt[fin], rs[fin] = rel(tau_g[fin], tau_c[fin]), rel(std_g[fin], std_c[fin])
    rej_g, rej_c = padj_g < 0.05, padj_c < 0.05
    out = {
        "config": args.config, "seed": args.seed, "n": n, "p": p, "a": a, "r": r, "family": fam,
        "cpu_side": "unmodified reference package (oracle/_ref), statsmodels served by oracle/shims",
        "gpu_seconds": round(t_gpu, 3), "reference_seconds": round(st["t_total"], 1),
        "reference_phases_s": {k: round(v, 1) for k, v in st.items() if k.startswith("t_")},
        "reference_cores": os.cpu_count(), "reference_jit_warmup_s_excluded": round(jit, 1),
        "reference_updates_per_s": st["cell_gene_updates"] / st["t_total"],
        "iters_gpu": [len(r1["hist"]) - 1, len(r2["hist"]) - 1], "iters_ref": [len(q1["hist"]) - 1, len(q2["hist"]) - 1],
        "irls_gene_iters_ref": st["irls_gene_iters"],
        "irls_gene_iters_gpu": int(sum(int(f["n_iter"].sum()) for f in fits)),
        "pairs_compared": int(fin.sum()), "pairs_total": int(a * p),
        "same_nonestimable_pattern": bool(np.array_equal(np.isfinite(std_g), np.isfinite(std_c))),
        "tau_rel_max": float(rt.max()), "tau_rel_p999": float(np.quantile(rt[fin], 0.999)),
        "tau_rel_median": float(np.median(rt[fin])),
        "std_rel_max": float(rs.max()), "std_rel_p999": float(np.quantile(rs[fin], 0.999)),
        "std_rel_median": float(np.median(rs[fin])),
        "tau_abs_max": float(np.abs(tau_g[fin] - tau_c[fin]).max()),
        "pairs_tau_above_1e-6": int((rt > 1e-6).sum()), "pairs_std_above_1e-6": int((rs > 1e-6).sum()),
        "pairs_above_1e-8": int(((rt > 1e-8) | (rs > 1e-8)).sum()),
        "rejections_gpu": int(rej_g.sum()), "rejections_ref": int(rej_c.sum()), "rejections_differ": int((rej_g != rej_c).sum()),
    }
    h1g, h1c, h2g, h2c = (np.asarray(x, dtype=float) for x in (r1["hist"], q1["hist"], r2["hist"], q2["hist"]))
    if len(h1g) == len(h1c) and len(h2g) == len(h2c):
        out["hist_rel_max"] = [float(np.max(np.abs(h1g - h1c) / np.abs(h1c))), float(np.max(np.abs(h2g - h2c) / np.abs(h2c)))]
    Ug, Uc = r2["U"], q2["U"]
    if Ug.shape == Uc.shape:
        out["U_max_abs_diff"] = float(np.abs(Ug - Uc).max())
        out["B_max_abs_diff"] = float(np.abs(r2["B_Gamma"] - q2["B_Gamma"]).max())

    # ---- stop-rule analysis of the genes behind the largest differences ----
    worst = ((rt > 1e-8) | (rs > 1e-8)).any(axis=0)
    wg = np.where(worst)[0]
    order = np.argsort(-np.maximum(rt, rs).max(axis=0)[wg])
    wg = wg[order][:args.max_probe]
    out["genes_with_pairs_above_1e-8"] = int(worst.sum())
    probes = []
    if len(wg):
        Yf = np.asarray(Y, dtype=np.float64)
        one = GlmDevice(n, 1)
        for j in wg:
            rec = {"gene": int(j), "tau_rel_max": float(rt[:, j].max()), "std_rel_max": float(rs[:, j].max()), "fits": {}}
            one.load_counts(np.ascontiguousarray(Yf[:, j:j + 1]))
            for name, f in zip(fit_names, fits):
                alpha = None
                if f["family"] == "nb" and f["nu"] is not None and not (f["nu"][j] > 100.0):
                    alpha = 1.0 / f["nu"][j]
                # CPU: the statsmodels restatement with its deviance history
                hist_c = []
                y = Yf[:, j]
                mu = (y + y.mean()) / 2.0
                o = np.zeros(n) if f["offset"] is None else f["offset"]
                eta = np.log(mu)
                hist_c.append(float(np.sum(og.unit_deviance(y, mu, alpha))))
                itc = 0
                try:
                    with np.errstate(all="ignore"):
                        for itc in range(1, min(f["maxiter"], 60) + 1):
                            w_ = mu if alpha is None else mu / (1.0 + alpha * mu)
                            z = eta + (y - mu) / mu - o
                            sw = np.sqrt(w_)
                            bta = np.linalg.lstsq(sw[:, None] * f["X"], sw * z, rcond=-1)[0]
                            eta = f["X"] @ bta + o
                            mu = np.exp(eta)
                            hist_c.append(float(np.sum(og.unit_deviance(y, mu, alpha))))
                            if not np.isfinite(hist_c[-1]) or abs(hist_c[-1] - hist_c[-2]) <= 1e-8:
                                break
                except Exception:  # noqa: BLE001  (a diverging fit: the guard / lasso path takes it)
                    itc = -1
                # device: deviance after m updates = a single-gene fit capped at m iterations
                itg = int(f["n_iter"][j])
                hist_g = []
                for m in range(max(1, itg - 2), itg + 2):
                    nu1 = None if f["nu"] is None else f["nu"][j:j + 1]
                    _, dv, it1, _ = one.fit(f["family"], f["X"], f["offset"], nu1, maxiter=m, d_guard=-1)
                    hist_g.append([m, float(dv[0]), int(it1[0])])
                rec["fits"][name] = {
                    "iters_device": itg, "iters_cpu": itc, "status_device": int(f["status"][j]),
                    "cpu_last_dev_changes": [abs(hist_c[i] - hist_c[i - 1]) for i in range(max(1, len(hist_c) - 3), len(hist_c))],
                    "device_dev_after_m_updates": hist_g,
                    "device_dev_changes": [abs(hist_g[i][1] - hist_g[i - 1][1]) for i in range(1, len(hist_g))]}
            probes.append(rec)
    out["probes"] = probes
    annotate(out)
    ws.close()
    line = json.dumps(out)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(line + "\n")
    print(line)


if __name__ == "__main__":
    main()
