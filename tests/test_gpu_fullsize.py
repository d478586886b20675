"""Full-size checks on the benchmark shapes (BASELINE.json configs 1 and 3) through properties that do
not need a CPU run of the same size: estimating-equation residuals of the fitted GLMs, monotone
descent and determinism of the alternating iteration, gene-permutation equivariance, and -- at the
CPU-runnable config-1 shape -- the C restatement of one alternating iteration itself."""
import numpy as np
import pytest
import scipy.linalg

from oracle import altmin as om

pytestmark = pytest.mark.gpu


def _batch(seed=3, p=8000, n_ctrl=2000, n_perts=15, cpp=133, r=10, fam="nb"):
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=n_ctrl, n_perts=n_perts, cells_per_pert=cpp, p=p, r=r, d_cov=1, family=fam, seed=seed,
                          base_lo=-1.5, base_hi=1.5)
    return sim


def test_glm_score_equations_full_size():
    """At the IRLS fixed point the estimating equations hold: X^T (y - mu) = 0 (Poisson, log link)
    and X^T [(y - mu) / (1 + mu / nu)] = 0 (NB2 with fixed size nu) -- checked for all 8000 genes
    of a benchmark-shaped batch, relative to the scale X^T (y + mu)."""
    from causarray_b200.device import GlmDevice
    sim = _batch()
    Y, A, U = sim["Y"], sim["A"], sim["U"]
    n, p = Y.shape
    X = np.c_[np.ones(n), U * 0.3, A]
    off = np.log(sim["size_scale"])
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    for fam, nu in (("poisson", None), ("nb", sim["disp"])):
        B, dv, n_iter, status = dev.fit(fam, X, off, nu, maxiter=100, tol=1e-8, d_guard=-1)
        ok = (status == 0) & (n_iter < 100)
        assert ok.mean() > 0.99
        mu = dev.fitted()
        np.testing.assert_allclose(mu[:, ok], np.exp(X @ B[ok].T + off[:, None]), rtol=1e-9)
        wres = (Y - mu) if fam == "poisson" else (Y - mu) / (1.0 + mu / nu[None, :])
        score = X.T @ wres                      # (kx, p)
        scale = np.abs(X).T @ (Y + mu) + 1.0
        rel = np.abs(score[:, ok]) / scale[:, ok]
        print(fam, "max relative score", rel.max(), "median iterations", np.median(n_iter))
        assert rel.max() < 2e-5                 # |dev - dev_prev| <= 1e-8 stops a step short of the exact root
        assert np.median(rel) < 1e-8


def test_alternating_iteration_descent_determinism_permutation():
    """Benchmark-shaped batch (3995 x 8000, k = 26), stage 1: the objective decreases monotonically,
    two runs agree bit for bit, and permuting the genes permutes B and leaves A unchanged up to
    summation-order rounding."""
    from causarray_b200.device import GcateDevice
    sim = _batch(seed=5)
    Y, Atr = sim["Y"], sim["A"]
    n, p = Y.shape
    rng = np.random.default_rng(0)
    X = np.c_[np.ones(n), Atr]
    d, r = X.shape[1], 10
    sf = sim["size_scale"] / np.exp(np.mean(np.log(sim["size_scale"])))
    A0 = np.c_[X, rng.standard_normal((n, r)) * 0.05]
    B0 = np.c_[np.log(Y.mean(0) + 0.05)[:, None], rng.standard_normal((p, d - 1 + r)) * 0.05]
    Q1 = scipy.linalg.qr(X, mode="economic")[0]

    def run(perm=None):
        Yp, B0p, nu = (Y, B0, sim["disp"]) if perm is None else (Y[:, perm], B0[perm], sim["disp"][perm])
        dev = GcateDevice(n, p, d, r, "nb", 100.0)
        dev.load_counts(np.ascontiguousarray(Yp), sf, nu)
        dev.set_factors(A0, np.ascontiguousarray(B0p))
        dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
        fs = [dev.nll()]
        for _ in range(4):
            fs.append(dev.update(1e3, 0.1, 0.5, 20, 1e-4))
        A, B = dev.get_factors()
        dev.close()
        return np.array(fs), A, B

    f1, A1, B1 = run()
    f2, A2, B2 = run()
    assert np.all(np.diff(f1) < 0), f1
    np.testing.assert_array_equal(f1, f2)
    np.testing.assert_array_equal(A1, A2)
    np.testing.assert_array_equal(B1, B2)
    perm = rng.permutation(p)
    f3, A3, B3 = run(perm)
    np.testing.assert_allclose(f3, f1, rtol=1e-11)
    np.testing.assert_allclose(B3, B1[perm], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(A3, A1, rtol=1e-7, atol=1e-9)


def test_config1_shape_vs_c_oracle():
    """BASELINE config 1 (n = 1000 x p = 2000 Poisson, r = 2, one treatment): two alternating
    iterations of stage 1 and two of stage 2 against the C restatement of update_with_mask."""
    from causarray_b200.device import GcateDevice
    from oracle import cport
    sim = _batch(seed=9, p=2000, n_ctrl=500, n_perts=1, cpp=500, r=2, fam="poisson")
    Y, Atr = sim["Y"], sim["A"]
    n, p = Y.shape
    rng = np.random.default_rng(1)
    X = np.c_[np.ones(n), rng.standard_normal((n, 2)) * 0.3, Atr]
    d, r, a = X.shape[1], 2, 1
    sf = sim["size_scale"] / np.exp(np.mean(np.log(sim["size_scale"])))
    nu = np.ones((1, p))
    Yn = Y / sf[:, None]
    Ys = om.log_h_terms(Yn, nu, "poisson", 100.0)
    A0 = np.c_[X, rng.standard_normal((n, r)) * 0.1]
    B0 = np.c_[np.log(Yn.mean(0) + 0.05)[:, None], rng.standard_normal((p, d - 1 + r)) * 0.1]
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    dev = GcateDevice(n, p, d, r, "poisson", 100.0)
    dev.load_counts(Y, sf, None)
    dev.set_factors(A0, B0)
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    A, B = A0.copy(), B0.copy()
    ones_g, ones_c, ag = np.ones(p, bool), np.ones(n, bool), np.full(p, 0.1)
    for it in range(2):
        f_ref, A, B = cport.update_with_mask(Yn, A, B, d, np.zeros((p, d)), Q1, None, "poisson", nu, Ys, 100.0, d, 1e5,
                                             0.1, 0.5, 20, 1e-4, ones_g, ones_c, ag)
        f = dev.update(1e5, 0.1, 0.5, 20, 1e-4)
        assert abs(f - f_ref) <= 1e-10 * abs(f_ref)
    Ad, Bd = dev.get_factors()
    np.testing.assert_allclose(Ad, A, atol=1e-9)
    np.testing.assert_allclose(Bd, B, atol=1e-9)
    Q2 = scipy.linalg.qr(B[:, d:], mode="economic")[0]
    lam = 0.02 / (np.abs(B[:, d - a:d]) + 1e-6)
    dev.set_stage(P1=None, P2=Q2, lam=lam, a=a)
    for it in range(2):
        f_ref, A, B = cport.update_with_mask(Yn, A, B, d, lam, None, Q2, "poisson", nu, Ys, 100.0, a, 1e5, 0.1, 0.5, 20,
                                             1e-4, ones_g, ones_c, ag)
        f = dev.update(1e5, 0.1, 0.5, 20, 1e-4)
        assert abs(f - f_ref) <= 1e-10 * abs(f_ref)
    Ad, Bd = dev.get_factors()
    np.testing.assert_allclose(Ad, A, atol=1e-9)
    np.testing.assert_allclose(Bd, B, atol=1e-9)


def test_benchmark_cells_vs_oracle_pipeline():
    """All 3995 cells of a benchmark-shaped batch (15 perturbations, r = 10, k = 26) with a subset of the
    genes, `fit_gcate` + `LFC` against the whole CPU oracle pipeline.  With this many cells nearly every
    gene of the second initialisation fit takes the regularised fallback, whose minimiser must be reached
    tightly on both sides (objective-only stopping left the unit-norm factor columns 1e-4 apart and
    `tau` 1e-7 apart -- see DESIGN.md section 5); iteration counts, histories and rejections must agree."""
    import warnings
    import pandas as pd
    import causarray_b200 as cb
    from causarray_b200.prep import prep_causarray_data
    from oracle import pipeline as op, cport
    sim = _batch(seed=11, p=96)
    Y, A = sim["Y"], sim["A"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    n, p = Y.shape
    a, r = A.shape[1], 10
    es = dict(rel_tol=2e-4, max_iters=12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r1, r2 = cb.fit_gcate(Y, X, A, r, family="nb", disp_glm=sim["disp"], offset=True, kwargs_es_1=dict(es),
                              kwargs_es_2=dict(es))
        ws = r2.pop("_ws")
        off = np.log(r2["kwargs_glm"]["size_factor"])
        df, _ = cb.LFC(Y, np.c_[X, r2["U"]], pd.DataFrame(A), np.c_[X_A, r2["U"]], family="nb", offset=off,
                       usevar="unequal", _glm=ws.glm, _trusted=True)
        ws.close()
        cport.set_threads(4)
        g, l, st = op.gcate_lfc_one_batch_cpu(Y, X, A, X_A, r, family="nb", disp=sim["disp"], es1=dict(es), es2=dict(es),
                                              usevar="unequal", n_jobs=4)
    assert len(r1["hist"]) == len(g["hist1"]) and len(r2["hist"]) == len(g["hist2"])
    np.testing.assert_allclose(r1["hist"], g["hist1"], rtol=1e-11)
    np.testing.assert_allclose(r2["hist"], g["hist2"], rtol=1e-11)
    tau_g, std_g = df["tau"].to_numpy().reshape(a, p), df["std"].to_numpy().reshape(a, p)
    padj_g = df["padj"].to_numpy().reshape(a, p)
    tau_c, std_c = l["tau"], np.sqrt(l["var"])
    assert np.array_equal(np.isfinite(std_g), np.isfinite(std_c))
    fin = np.isfinite(std_c) & (tau_c != 0)
    assert fin.sum() > 0.8 * a * p
    rt = np.abs(tau_g[fin] - tau_c[fin]) / np.maximum(np.abs(tau_c[fin]), 1e-12)
    rs = np.abs(std_g[fin] - std_c[fin]) / std_c[fin]
    print("tau rel max", rt.max(), "median", np.median(rt), "std rel max", rs.max())
    # north_star tolerance: LFC estimates and standard errors within 1e-6 relative.  (The convergence test
    # of the per-gene IRLS sums per-cell unit deviances on both sides, so the stop decisions agree; the
    # full-size records against the real reference -- profiles/r02_config{1,2,3}_ref_parity.json -- show
    # 3e-10 at worst with identical IRLS iteration totals.)
    assert np.quantile(rt, 0.99) <= 1e-8 and np.quantile(rs, 0.99) <= 1e-8
    assert rt.max() <= 1e-6 and rs.max() <= 1e-6
    assert np.array_equal(padj_g < 0.05, l["padj"] < 0.05)
