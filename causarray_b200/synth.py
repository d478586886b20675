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
