"""GPU parity of the S1 seam (update_with_mask / alter_min) against the golden
vectors produced by the real reference and against the CPU oracle."""
import os

import numpy as np
import pytest

from oracle import altmin as om

pytestmark = pytest.mark.gpu


def _dev(g, fam):
    from causarray_b200.device import GcateDevice
    Y, sf, nu = g["Y"], g["sf"], g["nu"]
    n, p = Y.shape
    d, r = int(g["d"]), int(g["r"])
    dev = GcateDevice(n, p, d, r, fam, float(g["thres_disp"]))
    dev.load_counts(Y, sf, nu)
    return dev


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_nll_grad_golden(golden_dir, fam):
    g = np.load(os.path.join(golden_dir, "kernels_%s.npz" % fam))
    dev = _dev(g, fam)
    dev.set_factors(g["A0"], g["B0"])
    f = dev.nll()
    print(fam, "nll", f, g["nll"], f - g["nll"])
    assert abs(f - g["nll"]) <= 1e-11 * abs(g["nll"])
    gB = dev.grad(False)
    gA = dev.grad(True)
    print(" gradB", np.abs(gB - g["grad_B"]).max(), "gradA", np.abs(gA - g["grad_A"]).max())
    np.testing.assert_allclose(gB, g["grad_B"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gA, g["grad_A"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("fam", ["nb", "poisson", "nb_clip"])
def test_update_with_mask_golden(golden_dir, fam):
    """(nb_clip: the golden was produced with C = 0.01, i.e. a norm-ball radius of 0.02 that rescales most
    cell and gene gradients -- project_norm_ball, causarray/gcate_opt.py:10-26)"""
    g = np.load(os.path.join(golden_dir, "kernels_%s.npz" % fam))
    d, a, r = int(g["d"]), int(g["a"]), int(g["r"])
    C = float(g["C"]) if "C" in g else (1e3 if fam == "nb" else 1e5)
    fam = fam.split("_")[0]
    # stage 1 (P1), masks + per-gene alpha, three consecutive iterations
    dev = _dev(g, fam)
    dev.set_factors(g["A0"], g["B0"])
    dev.set_stage(P1=g["Q1"], P2=None, lam=None, a=d)
    dev.set_masks(g["gene_active"], g["cell_active"], g["alpha_gene"])
    for it in range(3):
        f = dev.update(C, 0.1, 0.5, 20, 1e-4)
        A, B = dev.get_factors()
        dA, dB = np.abs(A - g["s1_A_%d" % it]).max(), np.abs(B - g["s1_B_%d" % it]).max()
        print(fam, "s1", it, f - g["s1_f"][it], dA, dB)
        assert abs(f - g["s1_f"][it]) <= 1e-10 * abs(g["s1_f"][it])
        assert dA <= 1e-9 and dB <= 1e-9
    # stage 2 (P2 + L1)
    dev.set_factors(g["s2_A_in"], g["s2_B_in"])
    dev.set_stage(P1=None, P2=g["Q2"], lam=g["lam2"], a=a)
    dev.set_masks(g["gene_active"], g["cell_active"], g["alpha_gene"])
    for it in range(3):
        f = dev.update(C, 0.1, 0.5, 20, 1e-4)
        A, B = dev.get_factors()
        dA, dB = np.abs(A - g["s2_A_%d" % it]).max(), np.abs(B - g["s2_B_%d" % it]).max()
        print(fam, "s2", it, f - g["s2_f"][it], dA, dB)
        assert abs(f - g["s2_f"][it]) <= 1e-10 * abs(g["s2_f"][it])
        assert dA <= 1e-9 and dB <= 1e-9
    # neither P1 nor P2 (reference T4 configuration, tests/test_gcate_convergence.py:203-240)
    dev.set_factors(g["A0"], g["B0"])
    dev.set_stage(P1=None, P2=None, lam=np.zeros((g["Y"].shape[1], d)), a=d)
    dev.set_masks(None, None, None)
    f = dev.update(C, 0.1, 0.5, 20, 1e-4)
    A, B = dev.get_factors()
    print(fam, "s0", f - g["s0_f"], np.abs(A - g["s0_A"]).max(), np.abs(B - g["s0_B"]).max())
    assert abs(f - g["s0_f"]) <= 1e-10 * abs(g["s0_f"])
    np.testing.assert_allclose(A, g["s0_A"], atol=1e-8)
    np.testing.assert_allclose(B, g["s0_B"], atol=1e-8)


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_alter_min_golden(golden_dir, fam):
    """Whole two-stage optimiser vs the real reference (explicit A0/B0)."""
    import scipy.linalg
    from causarray_b200.device import GcateDevice
    g = np.load(os.path.join(golden_dir, "alter_min_%s.npz" % fam))
    Y, X = g["Y"], g["X"]
    n, p = Y.shape
    d, r, a = X.shape[1], int(g["r"]), int(g["a"])
    dev = GcateDevice(n, p, d, r, fam, 100.0)
    dev.load_counts(Y, g["sf"], g["disp"])
    ls = {"alpha": 0.1, "beta": 0.5, "max_iters": 20, "tol": 1e-4, "C": 1e3 if fam == "nb" else 1e5,
          "tol_gene": 1e-4, "tol_cell": 1e-4, "recheck_interval": 10, "sparsity_threshold": 0.5,
          "sparsity_boost": 2.0, "warmup_iters": 0}
    es = {"max_iters": 30, "warmup": 0, "patience": 5, "tolerance": 0.0, "rel_tol": 2e-4}
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    dev.set_factors(g["A0"], g["B0"])
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    res, hist = dev.alter_min(ls, es, 0.0, d)
    A, B = dev.get_factors()
    print(fam, "stage1 n_iter", res.n_iter, int(g["s1_n_iter"]), "hist diff",
          np.abs(hist - g["s1_hist"][:len(hist)]).max() if len(hist) == len(g["s1_hist"]) else "LEN MISMATCH",
          "dA", np.abs(A - g["s1_A"]).max(), "dB", np.abs(B - g["s1_B"]).max(),
          "updates", res.total_gene_updates, int(g["s1_gene_updates"]), res.total_cell_updates,
          int(g["s1_cell_updates"]), "evals", res.cell_evals, res.gene_evals)
    assert res.n_iter == int(g["s1_n_iter"])
    np.testing.assert_allclose(hist, g["s1_hist"], rtol=1e-9)
    np.testing.assert_allclose(A, g["s1_A"], atol=1e-7)
    np.testing.assert_allclose(B, g["s1_B"], atol=1e-7)
    assert res.total_gene_updates == int(g["s1_gene_updates"])
    assert res.total_cell_updates == int(g["s1_cell_updates"])
    # stage 2 from the reference's own stage-1 result
    dev.set_factors(g["s1_A"], g["s1_B"])
    dev.set_stage(P1=None, P2=g["Q2"], lam=None, a=a)
    res2, hist2 = dev.alter_min(ls, es, 0.05, a)
    A2, B2 = dev.get_factors()
    print(fam, "stage2 n_iter", res2.n_iter, int(g["s2_n_iter"]), "dA", np.abs(A2 - g["s2_A"]).max(), "dB",
          np.abs(B2 - g["s2_B"]).max())
    assert res2.n_iter == int(g["s2_n_iter"])
    np.testing.assert_allclose(hist2, g["s2_hist"], rtol=1e-9)
    np.testing.assert_allclose(A2, g["s2_A"], atol=1e-7)
    np.testing.assert_allclose(B2, g["s2_B"], atol=1e-7)


def test_oracle_vs_cuda_ragged():
    """Odd shapes (n, p not multiples of 8, k needing padding) against the CPU oracle."""
    from causarray_b200.device import GcateDevice
    from causarray_b200.synth import simulate_screen
    import scipy.linalg
    rng = np.random.default_rng(5)
    sim = simulate_screen(n_ctrl=37, n_perts=2, cells_per_pert=13, p=101, r=3, d_cov=2, family="nb", seed=5)
    Y = sim["Y"]
    X = np.c_[sim["X"], sim["A"]]
    n, p = Y.shape
    d, r = X.shape[1], 3
    sf = np.exp(rng.standard_normal(n) * 0.2)
    nu = sim["disp"].reshape(1, -1)
    A0 = np.c_[X, rng.standard_normal((n, r)) * 0.2]
    B0 = rng.standard_normal((p, d + r)) * 0.2
    Yn = Y / sf[:, None]
    Ys = om.log_h_terms(Yn, nu, "nb", 100.0)
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    A, B = A0.copy(), B0.copy()
    f_ref, A, B = om.update_with_mask(Yn, A, B, d, np.zeros((p, d)), Q1, None, "nb", nu, Ys, 100.0, d, 1e3, 0.1, 0.5,
                                      20, 1e-4, np.ones(p, bool), np.ones(n, bool), np.full(p, 0.1))
    dev = GcateDevice(n, p, d, r, "nb", 100.0)
    dev.load_counts(Y, sf, nu[0])
    dev.set_factors(A0, B0)
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
    Ad, Bd = dev.get_factors()
    print("ragged", f - f_ref, np.abs(Ad - A).max(), np.abs(Bd - B).max())
    assert abs(f - f_ref) <= 1e-10 * abs(f_ref)
    np.testing.assert_allclose(Ad, A, atol=1e-9)
    np.testing.assert_allclose(Bd, B, atol=1e-9)
    # drop-in seam with a pre-normalised Y and an explicit Ys matrix
    dev2 = GcateDevice(n, p, d, r, "nb", 100.0)
    dev2.load_normalized(Yn, Ys, nu[0])
    dev2.set_factors(A0, B0)
    dev2.set_stage(P1=Q1, P2=None, lam=None, a=d)
    f2 = dev2.update(1e3, 0.1, 0.5, 20, 1e-4)
    assert abs(f2 - f_ref) <= 1e-10 * abs(f_ref)
