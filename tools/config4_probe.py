"""BASELINE config 4 (n = 100 000 cells x p = 10 000 genes, 21 covariates, one treatment, r = 10; k = 32)
through the public API on ONE GPU: `fit_gcate` + `LFC(usevar='unequal')`.  The counts are drawn on the
device (gamma-Poisson, gene chunks) because numpy needs minutes for 1e9 negative-binomial draws; they
are then handed to the API as an ordinary host float64 array (8 GB).  Prints one JSON line with the
phase times, iteration counts and size-independent checks (monotone objective, finite factors,
recovery of the simulated effects, estimating equations of the outcome GLM on a gene sample).

    python tools/config4_probe.py [--n 100000] [--p 10000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def mem_available_gb():
    with open("/proc/meminfo") as f:
        for line in f:
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 2 ** 20
    return 0.0


def simulate(n, p, d_cov, r, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g, device="cuda", dtype=torch.float64)
    uni = lambda *s: torch.rand(*s, generator=g, device="cuda", dtype=torch.float64)
    Xc = rnd(n, d_cov)
    U = rnd(n, r)
    logit = 0.4 * U[:, 0] - 0.3 * U[:, 1] + 0.2 * Xc[:, 0]
    A = (uni(n) < torch.sigmoid(logit)).to(torch.float64)
    lib = torch.exp(0.3 * rnd(n))
    b0 = uni(p) * 1.5 + 0.2
    Bc = 0.05 * rnd(p, d_cov)
    Gm = 0.25 * rnd(p, r)
    tau = torch.zeros(p, device="cuda", dtype=torch.float64)
    nn = p // 10
    idx = torch.randperm(p, generator=g, device="cuda")[:nn]
    tau[idx] = (uni(nn) + 0.5) * torch.sign(rnd(nn))
    nu = 1.0 / (0.5 + 1.5 * uni(p))
    Y = np.empty((n, p), dtype=np.float64)
    step = 500
    for lo in range(0, p, step):
        hi = min(p, lo + step)
        eta = b0[lo:hi][None, :] + Xc @ Bc[lo:hi].T + U @ Gm[lo:hi].T + A[:, None] * tau[lo:hi][None, :]
        mu = torch.exp(eta) * lib[:, None]
        lam = torch._standard_gamma(nu[lo:hi][None, :].expand(n, -1).contiguous()) * (mu / nu[lo:hi][None, :])
        y = torch.poisson(lam, generator=g)
        y[0, y.sum(0) == 0] = 1.0
        Y[:, lo:hi] = y.cpu().numpy()
    out = dict(Y=Y, X=np.c_[np.ones(n), Xc.cpu().numpy()], A=A.cpu().numpy()[:, None], tau=tau.cpu().numpy(),
               nu=nu.cpu().numpy(), U=U.cpu().numpy())
    del Xc, U, Bc, Gm, eta, mu, lam, y
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--p", type=int, default=10000)
    ap.add_argument("--r", type=int, default=10)
    ap.add_argument("--dcov", type=int, default=20)
    args = ap.parse_args()
    need = args.n * args.p * 8 / 2 ** 30
    avail = mem_available_gb()
    print("host MemAvailable %.1f GB, counts %.1f GB" % (avail, need), file=sys.stderr)
    if avail < 3.5 * need:
        args.n = int(args.n * avail / (3.5 * need))
        print("reduced to n = %d" % args.n, file=sys.stderr)
    import causarray_b200 as cb
    from causarray_b200 import _timing
    t0 = time.time()
    sim = simulate(args.n, args.p, args.dcov, args.r)
    t_sim = time.time() - t0
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    n, p = Y.shape
    es = dict(rel_tol=2e-4, max_iters=30)
    torch.cuda.synchronize()
    _timing.enable(True)
    t0 = time.time()
    r1, r2 = cb.fit_gcate(Y, X, A, args.r, family="nb", disp_glm=sim["nu"], kwargs_es_1=es, kwargs_es_2=es)
    torch.cuda.synchronize()
    t_fit = time.time() - t0
    ws = r2.pop("_ws")
    W = np.c_[X, r2["U"]]
    t0 = time.time()
    df, est = cb.LFC(Y, W, A, W_A=W, family="nb", offset=np.log(np.ravel(r2["kwargs_glm"]["size_factor"])),
                     disp_glm=sim["nu"], usevar="unequal", _glm=ws.glm, _trusted=True)
    torch.cuda.synchronize()
    t_lfc = time.time() - t0
    # Calibration check (round-1 review: FDP 0.85 at power 1 went unexplained).  The same LFC with the TRUE
    # latent confounders as covariates, and the effect sizes behind the false discoveries: with n = 1e5 the
    # standard errors are ~5e-3, so a residual-confounding bias of 1e-2 in log fold change -- invisible
    # at the reference's test sizes -- already rejects; it comes from the estimated factors (the statistical
    # method), not from the arithmetic, if the oracle-U run is calibrated.
    df_or, _ = cb.LFC(Y, np.c_[X, sim["U"]], A, W_A=np.c_[X, sim["U"]], family="nb",
                      offset=np.log(np.ravel(r2["kwargs_glm"]["size_factor"])), disp_glm=sim["nu"], usevar="unequal",
                      _glm=ws.glm, _trusted=True)
    ws.close()
    h1, h2 = np.asarray(r1["hist"]), np.asarray(r2["hist"])
    tau_hat = df["tau"].to_numpy()
    nz = sim["tau"] != 0
    rej = df["padj"].to_numpy() < 0.05
    updates = n * p * (len(h1) - 1 + len(h2) - 1)
    out = {
        "config": "C4", "n": n, "p": p, "k": int(r2["X_U"].shape[1]), "t_simulate_s": round(t_sim, 1),
        "t_fit_gcate_s": round(t_fit, 2), "t_lfc_s": round(t_lfc, 2),
        "iters_stage1": len(h1) - 1, "iters_stage2": len(h2) - 1,
        "hist1_monotone": bool(np.all(np.diff(h1) < 0)), "hist2_first_last": [float(h2[0]), float(h2[-1])],
        "factors_finite": bool(np.isfinite(r2["X_U"]).all() and np.isfinite(r2["B_Gamma"]).all()),
        "altmin_updates_per_s_incl_init": updates / t_fit,
        "tau_corr_nonnull": float(np.corrcoef(tau_hat[nz], sim["tau"][nz])[0, 1]),
        "tau_rmse_nonnull": float(np.sqrt(np.mean((tau_hat[nz] - sim["tau"][nz]) ** 2))),
        "power": float(rej[nz].mean()), "false_discovery_prop": float(rej[~nz].sum() / max(1, rej.sum())),
        "n_nan_tau": int(np.isnan(tau_hat).sum()),
        "calibration": {
            "null_genes": int((~nz).sum()), "false_discoveries": int(rej[~nz].sum()),
            "abs_tau_false_discoveries_quantiles_50_90_99": [float(q) for q in np.quantile(np.abs(tau_hat[~nz][rej[~nz]]), [0.5, 0.9, 0.99])] if rej[~nz].any() else None,
            "abs_tau_all_nulls_quantiles_50_90_99": [float(q) for q in np.quantile(np.abs(tau_hat[~nz]), [0.5, 0.9, 0.99])],
            "median_std_nulls": float(np.nanmedian(df["std"].to_numpy()[~nz])),
            "null_tstat_median_and_mad_sd": [float(np.nanmedian(df["stat"].to_numpy()[~nz])),
                                             float(1.4826 * np.nanmedian(np.abs(df["stat"].to_numpy()[~nz] - np.nanmedian(df["stat"].to_numpy()[~nz]))))],
            "fdp_with_empirical_null_adjustment": float(((df["padj_emp_null_adj"].to_numpy() < 0.05) & ~nz).sum()
                                                         / max(1, (df["padj_emp_null_adj"].to_numpy() < 0.05).sum())),
            "oracle_U_covariates": {
                "false_discovery_prop": float(((df_or["padj"].to_numpy() < 0.05) & ~nz).sum() / max(1, (df_or["padj"].to_numpy() < 0.05).sum())),
                "power": float((df_or["padj"].to_numpy() < 0.05)[nz].mean()),
                "null_tstat_mad_sd": float(1.4826 * np.nanmedian(np.abs(df_or["stat"].to_numpy()[~nz] - np.nanmedian(df_or["stat"].to_numpy()[~nz]))))},
        },
        "phases": {k: round(v, 3) for k, v in sorted(_timing.TOTALS.items(), key=lambda kv: -kv[1])[:14]},
        "gpu_mem_peak_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
