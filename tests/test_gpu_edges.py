"""GPU edge cases of the three seams: empty masks, tiny / ragged shapes, degenerate designs,
wide designs, and the error behaviour of the C ABI (return codes, no exceptions across it)."""
import os
import numpy as np
import pytest
import scipy.linalg

from oracle import altmin as om
from oracle import glm as og

pytestmark = pytest.mark.gpu


def _problem(n_ctrl, n_perts, cpp, p, r, fam="nb", seed=0, d_cov=1):
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=n_ctrl, n_perts=n_perts, cells_per_pert=cpp, p=p, r=r, d_cov=d_cov, family=fam,
                          seed=seed, base_lo=0.0, base_hi=2.0)
    rng = np.random.default_rng(seed)
    Y = sim["Y"]
    X = np.c_[sim["X"], sim["A"]]
    n = Y.shape[0]
    d = X.shape[1]
    A0 = np.c_[X, rng.standard_normal((n, r)) * 0.2]
    B0 = rng.standard_normal((p, d + r)) * 0.2
    sf = np.exp(rng.standard_normal(n) * 0.2)
    return Y, X, A0, B0, sf, sim["disp"], d


def test_all_rows_inactive_is_identity():
    """Masks all False: neither factor moves and the value is the plain nll (gcate_opt.py:170, 194)."""
    from causarray_b200.device import GcateDevice
    Y, X, A0, B0, sf, nu, d = _problem(40, 2, 12, 53, 2)
    n, p = Y.shape
    dev = GcateDevice(n, p, d, 2, "nb", 100.0)
    dev.load_counts(Y, sf, nu)
    dev.set_factors(A0, B0)
    f0 = dev.nll()
    dev.set_stage(P1=None, P2=None, lam=np.zeros((p, d)), a=d)
    dev.set_masks(np.zeros(p, bool), np.zeros(n, bool), np.full(p, 0.1))
    f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
    A, B = dev.get_factors()
    np.testing.assert_array_equal(A, A0)
    np.testing.assert_array_equal(B, B0)
    assert abs(f - f0) <= 1e-12 * abs(f0)


@pytest.mark.parametrize("fam", ["nb", "poisson"])
@pytest.mark.parametrize("shape", [(7, 1, 2, 5, 1), (20, 1, 5, 9, 1), (130, 3, 9, 7, 2), (9, 2, 3, 260, 3)])
def test_tiny_and_ragged_shapes_vs_oracle(fam, shape):
    """Fewer rows than a tile, fewer columns than a chunk, one more than a chunk, p < 8."""
    from causarray_b200.device import GcateDevice
    n_ctrl, n_perts, cpp, p, r = shape
    Y, X, A0, B0, sf, nu, d = _problem(n_ctrl, n_perts, cpp, p, r, fam=fam, seed=3)
    n = Y.shape[0]
    nu = nu.reshape(1, -1) if fam == "nb" else np.ones((1, p))
    Yn = Y / sf[:, None]
    Ys = om.log_h_terms(Yn, nu, fam, 100.0)
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    C = 1e3 if fam == "nb" else 1e5
    A, B = A0.copy(), B0.copy()
    f_ref, A, B = om.update_with_mask(Yn, A, B, d, np.zeros((p, d)), Q1, None, fam, nu, Ys, 100.0, d, C, 0.1, 0.5, 20,
                                      1e-4, np.ones(p, bool), np.ones(n, bool), np.full(p, 0.1))
    dev = GcateDevice(n, p, d, r, fam, 100.0)
    dev.load_counts(Y, sf, nu[0] if fam == "nb" else None)
    dev.set_factors(A0, B0)
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    f = dev.update(C, 0.1, 0.5, 20, 1e-4)
    Ad, Bd = dev.get_factors()
    assert abs(f - f_ref) <= 1e-10 * abs(f_ref)
    np.testing.assert_allclose(Ad, A, atol=1e-9)
    np.testing.assert_allclose(Bd, B, atol=1e-9)


def test_stage2_backtracking_rows_vs_oracle():
    """Stage 2 with an L1 weight large enough that many genes fail the first candidate: the
    multi-candidate phase of the line search against the oracle's sequential walk, with per-gene
    initial steps so that genes stop after different numbers of candidates."""
    from causarray_b200.device import GcateDevice
    Y, X, A0, B0, sf, nu, d = _problem(60, 3, 15, 97, 2, seed=11)
    n, p = Y.shape
    a, r = 3, 2
    nu = nu.reshape(1, -1)
    Yn = Y / sf[:, None]
    Ys = om.log_h_terms(Yn, nu, "nb", 100.0)
    Q2 = scipy.linalg.qr(B0[:, d:], mode="economic")[0]
    rng = np.random.default_rng(2)
    lam = 0.5 / (np.abs(B0[:, d - a:d]) + 1e-6)
    alpha_gene = np.where(rng.random(p) < 0.5, 0.1, 0.27)
    dev = GcateDevice(n, p, d, r, "nb", 100.0)
    dev.load_counts(Y, sf, nu[0])
    dev.set_factors(A0, B0)
    dev.set_stage(P1=None, P2=Q2, lam=lam, a=a)
    dev.set_masks(np.ones(p, bool), np.ones(n, bool), alpha_gene)
    A, B = A0.copy(), B0.copy()
    for it in range(4):
        f_ref, A, B = om.update_with_mask(Yn, A, B, d, lam, None, Q2, "nb", nu, Ys, 100.0, a, 1e3, 0.1, 0.5, 20, 1e-4,
                                          np.ones(p, bool), np.ones(n, bool), alpha_gene)
        f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
        Ad, Bd = dev.get_factors()
        assert abs(f - f_ref) <= 1e-10 * abs(f_ref), (it, f, f_ref)
        np.testing.assert_allclose(Ad, A, atol=1e-9)
        np.testing.assert_allclose(Bd, B, atol=1e-9)


def test_glm_intercept_only_is_log_mean():
    """kx = 1 without offset: the Poisson MLE is log(mean y) (one tile, first / regular / finalize passes)."""
    from causarray_b200.device import GlmDevice
    rng = np.random.default_rng(0)
    n, p = 333, 41
    Y = rng.poisson(rng.uniform(0.5, 20.0, p)[None, :], size=(n, p)).astype(float)
    Y[:, 0] += 1.0
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    B, dv, n_iter, status = dev.fit("poisson", np.ones((n, 1)), None, None, maxiter=100, tol=1e-8, d_guard=1)
    assert (status == 0).all()
    np.testing.assert_allclose(B[:, 0], np.log(Y.mean(0)), rtol=1e-10)
    np.testing.assert_allclose(dev.fitted(), np.broadcast_to(Y.mean(0), (n, p)), rtol=1e-9)


def test_glm_wide_design_vs_oracle():
    """kx = 41 (six tiles, one CTA per SM variant) against the IRLS oracle."""
    from causarray_b200.device import GlmDevice
    rng = np.random.default_rng(4)
    n, p, kx = 500, 24, 41
    X = np.c_[np.ones(n), rng.standard_normal((n, kx - 1)) * 0.3]
    beta = np.c_[rng.uniform(0.5, 2.0, p), rng.standard_normal((p, kx - 1)) * 0.2]
    Y = rng.poisson(np.exp(X @ beta.T)).astype(float)
    off = rng.standard_normal(n) * 0.1
    B_ref, _, _, _, _, status_ref, it_ref = og.fit_glm_original(Y, X, None, family="poisson", offset=off, maxiter=100)
    dev = GlmDevice(n, p)
    dev.load_counts(Y)
    B, dv, n_iter, status = dev.fit("poisson", X, off, None, maxiter=100, tol=1e-8, d_guard=-1)
    ok = (status == 0) & (status_ref == 0)
    assert ok.sum() >= p - 1
    np.testing.assert_allclose(B[ok], B_ref[ok], rtol=1e-7, atol=1e-8)


def test_abi_error_codes():
    """Invalid arguments come back as return codes with a message; nothing throws across the ABI."""
    import ctypes
    from causarray_b200 import _lib as L
    from causarray_b200.device import GcateDevice, GlmDevice
    with pytest.raises(ValueError, match="Family not recognized"):
        GcateDevice(10, 10, 2, 1, "gaussian")
    lib = L.load()
    h = ctypes.c_void_p()
    rc = lib.ca_gcate_create(ctypes.byref(h), 10, 10, 2, 1, 7, ctypes.c_double(100.0))
    assert rc == -2 and b"Family not recognized" in lib.ca_last_error()
    rc = lib.ca_gcate_create(ctypes.byref(h), 0, 10, 2, 1, 0, ctypes.c_double(100.0))
    assert rc == -2
    dev = GcateDevice(12, 9, 2, 1, "poisson")
    with pytest.raises((ValueError, RuntimeError)):
        dev.update(1e3, 0.1, 0.5, 20, 1e-4)          # no counts / factors loaded
    g = GlmDevice(12, 9)
    with pytest.raises((ValueError, RuntimeError)):
        g.fit("poisson", np.ones((12, 1)))            # no counts loaded
    g.load_counts(np.ones((12, 9)))
    with pytest.raises((ValueError, RuntimeError)):
        g.fit("poisson", np.ones((12, 70)))           # design wider than 64 columns


def test_aipw_single_cell_arm_gives_nan_variance():
    """Welch variance with one treated cell: var / df are NaN for that treatment (DR_learner.py:440-470)."""
    from causarray_b200.device import aipw_lfc
    rng = np.random.default_rng(1)
    n, p, a = 60, 17, 2
    Y = rng.poisson(3.0, size=(n, p)).astype(float)
    A = np.zeros((n, a))
    A[:1, 0] = 1            # a single treated cell
    A[1:21, 1] = 1
    W = np.ones((n, 1))
    Bcov = np.log(Y.mean(0))[:, None]
    Btrt = np.zeros((p, a))
    pi = np.full((n, a), 0.3)
    out = aipw_lfc(Y, W, Bcov, Btrt, np.zeros(n), np.ones(n), A, pi, usevar="unequal")
    assert (np.isnan(out["var"][0]) | np.isinf(out["var"][0])).all()      # NaN, or inf where the expression filter applies
    assert np.isfinite(out["var"][1]).mean() > 0.8      # (the low-expression / small-difference filter sets inf)


def test_fit_gcate_input_validation():
    """_check_input (causarray/gcate.py:8-51): non-integer counts are rounded with a warning, an
    all-zero gene and a near-singular covariate matrix raise the reference's ValueErrors."""
    import warnings
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=80, n_perts=2, cells_per_pert=20, p=40, r=1, seed=2, base_lo=0.5, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    kw = dict(family="nb", disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=3), kwargs_es_2=dict(max_iters=3))
    Yz = Y.copy()
    Yz[:, 5] = 0.0
    with pytest.raises(ValueError, match="non-expressed"):
        cb.fit_gcate(Yz, X, A, 1, **kw)
    Xs = np.c_[X, X[:, 0] * (1 + 1e-9)]
    with pytest.raises(ValueError, match="near singular"):
        cb.fit_gcate(Y.copy(), Xs, A, 1, **kw)
    with pytest.raises(ValueError, match="ndim"):
        cb.fit_gcate(Y[:, 0], X, A, 1, **kw)
    Yf = Y + 0.3
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        r1, r2 = cb.fit_gcate(Yf, X, A, 1, **kw)
    assert any("not integer-valued" in str(w.message) for w in rec)
    r2.pop("_ws").close()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q1, q2 = cb.fit_gcate(np.round(Yf), X, A, 1, **kw)
    q2.pop("_ws").close()
    np.testing.assert_array_equal(r2["B_Gamma"], q2["B_Gamma"])


@pytest.mark.gpu
def test_lfc_with_fdx_control_golden():
    """LFC(fdx=True) through the dense influence-value path against the REAL reference
    (tests/golden/make_golden.py::golden_fdx): same seeded bootstrap stream, identical ``rej`` column."""
    import warnings
    import causarray_b200 as cb
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fdx.npz"))
    np.random.seed(321)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df, est = cb.LFC(g["Y"].copy(), g["W"], g["A"], W_A=g["W"], family="nb", offset=g["offset"],
                         Y_hat=g["Y_hat"].copy(), pi_hat=g["pi"].copy(), usevar="unequal", fdx=True, fdx_alpha=0.1,
                         fdx_c=0.2)
    np.testing.assert_array_equal(df["rej"].to_numpy().astype(float), g["lfc_rej"])
    np.testing.assert_allclose(df["tau"].to_numpy().astype(float), g["lfc_tau"], rtol=1e-9, atol=1e-12)
    a, b = df["padj"].to_numpy().astype(float), g["lfc_padj"]
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-12, equal_nan=True)
