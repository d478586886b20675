"""Seeded synthetic count screens of the shapes named in BASELINE.json.

The data-generating process follows SURVEY.md section 8(d): negative-binomial (or
Poisson) counts with ``r`` latent confounders, one-hot perturbation
assignment against a shared control pool, per-cell library-size variation
and 10 % non-null genes per perturbation with ``|tau| ~ U(0.5, 1.5)``.  It is
modelled on the simulators of the reference's own tests
(``tests/test_inference_comprehensive.py:28-73`` ``_sim_nb``,
``tests/test_batch_fitting.py:15-45`` ``_make_data``) but is not a copy of
either: assignment is one-hot over ``a`` perturbations and confounded through
a per-perturbation shift of the latent factors.
"""
import numpy as np


def simulate_screen(n_ctrl=200, n_perts=4, cells_per_pert=60, p=300, r=2, d_cov=1,
                    family="nb", seed=0, base_lo=-1.5, base_hi=1.5, gamma_sd=0.3,
                    frac_nonnull=0.1, libsize_sd=0.3, poisson_cells=False):
    """Return ``dict(Y, X, A, U, Gamma, tau, disp, size_scale)``.

    ``Y (n, p)`` float64 counts, ``X (n, d_cov)`` with an intercept column first,
    ``A (n, n_perts)`` one-hot (control rows all zero), ``disp (p,)`` NB size
    parameters (``nu = 1/phi``, ``phi ~ U(0.5, 2)``).
    """
    rng = np.random.default_rng(seed)
    if poisson_cells:
        sizes = np.maximum(rng.poisson(cells_per_pert, n_perts), 2)
    else:
        sizes = np.full(n_perts, cells_per_pert)
    n = int(n_ctrl + sizes.sum())
    label = np.full(n, -1)
    pos = n_ctrl
    for j, s in enumerate(sizes):
        label[pos:pos + s] = j
        pos += s
    perm = rng.permutation(n)
    label = label[perm]
    A = np.zeros((n, n_perts))
    A[np.where(label >= 0)[0], label[label >= 0]] = 1.0

    X = np.ones((n, 1))
    if d_cov > 1:
        X = np.c_[X, rng.standard_normal((n, d_cov - 1))]
    U = rng.standard_normal((n, r))
    if r > 0:
        shift = rng.standard_normal((n_perts, r)) * 0.4
        U[label >= 0] += shift[label[label >= 0]]
    Gamma = rng.standard_normal((p, r)) * gamma_sd
    beta0 = rng.uniform(base_lo, base_hi, p)
    Bx = np.c_[beta0, rng.standard_normal((p, d_cov - 1)) * 0.2] if d_cov > 1 else beta0[:, None]
    tau = np.zeros((n_perts, p))
    k_nn = int(round(frac_nonnull * p))
    for j in range(n_perts):
        idx = rng.choice(p, k_nn, replace=False)
        tau[j, idx] = rng.choice([-1.0, 1.0], k_nn) * rng.uniform(0.5, 1.5, k_nn)
    size_scale = np.exp(rng.standard_normal(n) * libsize_sd)
    eta = X @ Bx.T + U @ Gamma.T + A @ tau + np.log(size_scale)[:, None]
    mu = np.exp(np.clip(eta, -10.0, 8.0))
    phi = rng.uniform(0.5, 2.0, p)
    disp = 1.0 / phi
    if family == "nb":
        Y = rng.negative_binomial(disp[None, :], disp[None, :] / (disp[None, :] + mu)).astype(np.float64)
    elif family == "poisson":
        Y = rng.poisson(mu).astype(np.float64)
    else:
        raise ValueError("Family not recognized")
    empty = np.where(~np.any(Y != 0, axis=0))[0]
    if empty.size:
        Y[rng.integers(0, n, empty.size), empty] = 1.0
    return dict(Y=Y, X=X, A=A, U=U, Gamma=Gamma, tau=tau, disp=disp, size_scale=size_scale, label=label)


def simulate_screen_fast(n_ctrl=2000, n_perts=200, cells_per_pert=390, p=8000, r=10, family="nb", seed=0,
                         base_lo=-1.5, base_hi=1.5, gamma_sd=0.3, frac_nonnull=0.1, libsize_sd=0.3,
                         counts_dtype=np.int32, device=None):
    """Whole-screen variant of :func:`simulate_screen` for the full ``gcate_lfc_batch`` measurement
    (BASELINE config 3: 2,000 controls + 200 perturbations x Poisson(390) cells x 8,000 genes = 6.4e8
    draws).  Same data-generating process; the small pieces (labels, covariates, loadings, effects)
    are drawn with numpy, the (n, p) gamma-Poisson draws with torch on ``device`` when a GPU is present
    (numpy's negative-binomial sampler needs ~2 minutes for this size).  Returns ``Y`` as
    ``counts_dtype`` host array, ``A`` one-hot float64, ``disp``.  Synthetic input only -- not part of
    the measured path."""
    import torch
    rng = np.random.default_rng(seed)
    sizes = np.maximum(rng.poisson(cells_per_pert, n_perts), 2)
    n = int(n_ctrl + sizes.sum())
    label = np.full(n, -1)
    pos = n_ctrl
    for j, sz in enumerate(sizes):
        label[pos:pos + sz] = j
        pos += sz
    label = label[rng.permutation(n)]
    U = rng.standard_normal((n, r))
    shift = rng.standard_normal((n_perts, r)) * 0.4
    U[label >= 0] += shift[label[label >= 0]]
    Gamma = rng.standard_normal((p, r)) * gamma_sd
    beta0 = rng.uniform(base_lo, base_hi, p)
    tau = np.zeros((n_perts, p))
    k_nn = int(round(frac_nonnull * p))
    for j in range(n_perts):
        idx = rng.choice(p, k_nn, replace=False)
        tau[j, idx] = rng.choice([-1.0, 1.0], k_nn) * rng.uniform(0.5, 1.5, k_nn)
    size_scale = np.exp(rng.standard_normal(n) * libsize_sd)
    disp = 1.0 / rng.uniform(0.5, 2.0, p)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    Y = np.empty((n, p), dtype=counts_dtype)
    tau_ext = np.vstack([tau, np.zeros((1, p))])          # label -1 -> the zero row
    tU, tG = torch.as_tensor(U, device=device), torch.as_tensor(Gamma, device=device)
    tb, tt = torch.as_tensor(beta0, device=device), torch.as_tensor(tau_ext, device=device)
    tl, ts = torch.as_tensor(label, device=device), torch.as_tensor(np.log(size_scale), device=device)
    td = torch.as_tensor(disp, device=device)
    step = 8192
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        eta = tU[lo:hi] @ tG.T + tb[None, :] + tt[tl[lo:hi]] + ts[lo:hi, None]
        mu = torch.exp(torch.clamp(eta, -10.0, 8.0))
        if family == "nb":
            lam = torch._standard_gamma(td[None, :].expand(hi - lo, p).contiguous(), generator=gen) * (mu / td[None, :])
        elif family == "poisson":
            lam = mu
        else:
            raise ValueError("Family not recognized")
        Y[lo:hi] = torch.poisson(lam, generator=gen).to(torch.int32).cpu().numpy().astype(counts_dtype, copy=False)
    empty = np.where(~np.any(Y != 0, axis=0))[0]
    if empty.size:
        Y[rng.integers(0, n, empty.size), empty] = 1
    A = np.zeros((n, n_perts))
    A[np.where(label >= 0)[0], label[label >= 0]] = 1.0
    return dict(Y=Y, A=A, disp=disp, label=label, tau=tau, size_scale=size_scale)
