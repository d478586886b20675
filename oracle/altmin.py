"""CPU (numpy) restatement of the GCATE alternating minimisation.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pinned against the real
reference by ``tests/golden/make_golden.py`` / ``tests/test_oracle_golden.py``.

Reference lines followed:

* ``nll``   ``causarray/gcate_likelihood.py:46-94``
* ``grad``  ``causarray/gcate_likelihood.py:98-142``
* ``project_norm_ball`` ``causarray/gcate_opt.py:10-26``
* ``line_search``       ``causarray/gcate_opt.py:30-82``
* ``update_with_mask``  ``causarray/gcate_opt.py:148-206``
* ``alter_min`` loop    ``causarray/gcate_opt.py:343-458``
* ``Early_Stopping``    ``causarray/utils.py:101-161``

The restatement is row-vectorised: every active row runs the same sequential
backtracking loop as the reference, a boolean ``searching`` mask stands in for
the per-row ``return``.
"""
import numpy as np
from scipy.special import gammaln


def entry_terms(Theta, Y, nu, family, thres_disp):
    """Per-entry natural parameter and cumulant (``gcate_likelihood.py:78-88``).

    Returns ``(theta, b)`` such that the log-likelihood kernel is ``Y*theta - b``.
    ``nu`` broadcasts against ``Theta``.
    """
    xi = np.minimum(Theta, 100.0)
    with np.errstate(over="ignore"):
        m = np.exp(xi)
    if family == "poisson":
        return xi, m
    if family != "nb":
        raise ValueError("Family not recognized")
    q = np.clip(1.0 / (1.0 + m / nu), 1e-6, 1 - 1e-6)
    pois = nu > thres_disp
    theta = np.where(pois, xi, np.log1p(-q))
    b = np.where(pois, m, -nu * np.log(q))
    return theta, b


def nll(Y, A, B, family, nu, Tys, thres_disp):
    """``-sum(Y*theta - b + Tys) / len(Y)`` with ``Theta = A @ B.T``."""
    Theta = A @ B.T
    theta, b = entry_terms(Theta, Y, nu, family, thres_disp)
    return -np.sum(Y * theta - b + Tys) / float(Y.shape[0])


def dldeta(Theta, Y, nu, family, thres_disp):
    """``-(Y - e^xi) * (q or 1)`` (``gcate_likelihood.py:126-138``)."""
    xi = np.minimum(Theta, 100.0)
    with np.errstate(over="ignore"):
        m = np.exp(xi)
    if family == "poisson":
        return -(Y - m)
    q = np.clip(1.0 / (1.0 + m / nu), 1e-6, 1 - 1e-6)
    return -(Y - m) * np.where(nu > thres_disp, 1.0, q)


def grad(Y, A, B, family, nu, thres_disp):
    """Gradient wrt ``B``: ``dldeta(A B^T).T @ A / len(Y)`` -> ``(B.shape[0], k)``."""
    G = dldeta(A @ B.T, Y, nu, family, thres_disp)
    return G.T @ A / float(Y.shape[0])


def _ball(g, radius):
    """Row-wise inf-norm ball projection (``gcate_opt.py:10-26``)."""
    nrm = np.max(np.abs(g), axis=1)
    scale = np.where(nrm > radius, radius / np.where(nrm > 0, nrm, 1.0), 1.0)
    return g * scale[:, None]


def _soft(x, lam):
    return np.sign(x) * np.maximum(np.abs(x) - lam, 0.0)


def line_search_rows(Yr, M, X0, G, d, family, nu, Tys_r, thres_disp, start, lam,
                     alpha, beta, max_iters, tol, nu_per_row):
    """Backtracking line search for a block of rows at once.

    ``Yr (m, c)`` rows of the response, ``M (c, k)`` the fixed factor, ``X0, G
    (m, k)`` current rows and (projected) gradients, ``lam (m, d-start)`` or 0,
    ``alpha (m,)`` initial steps.  ``nu`` is ``(c,)`` (per column) when
    ``nu_per_row`` is False and ``(m,)`` (one per row) otherwise.  Row
    normaliser is ``c`` (``nll`` divides by ``len(Y)``, ``gcate_opt.py:51``).

    Returns ``(X1, n_eval)``.
    """
    m, c = Yr.shape
    nu_b = nu[:, None] if nu_per_row else nu[None, :]
    lam = np.broadcast_to(np.asarray(lam, dtype=np.float64), (m, d - start)) if d > start else np.zeros((m, 0))

    def fval(X):
        theta, b = entry_terms(X @ M.T, Yr, nu_b, family, thres_disp)
        f = -np.sum(Yr * theta - b + Tys_r, axis=1) / float(c)
        if d > start:
            f = f + np.mean(lam * np.abs(X[:, start:d]), axis=1)
        return f

    f0 = fval(X0)
    norm_g = np.sqrt(np.sum(G * G, axis=1))
    t = alpha.astype(np.float64).copy()
    X1 = X0.copy()
    searching = np.ones(m, dtype=bool)
    f01 = np.zeros(m)
    f1 = np.zeros(m)
    n_eval = np.zeros(m, dtype=np.int64)
    for i in range(max_iters):
        idx = np.where(searching)[0]
        if idx.size == 0:
            break
        Xc = X0[idx] - t[idx, None] * G[idx]
        if d > start:
            Xc[:, start:d] = _soft(Xc[:, start:d], lam[idx])
        theta, b = entry_terms(Xc @ M.T, Yr[idx], (nu[idx, None] if nu_per_row else nu_b), family, thres_disp)
        fc = -np.sum(Yr[idx] * theta - b + Tys_r[idx], axis=1) / float(c)
        if d > start:
            fc = fc + np.mean(lam[idx] * np.abs(Xc[:, start:d]), axis=1)
        n_eval[idx] += 1
        X1[idx] = Xc
        f1[idx] = fc
        if i == 0:
            f01[idx] = fc
        ok = fc < f0[idx] - tol * t[idx] * norm_g[idx]
        searching[idx[ok]] = False
        rest = idx[~ok]
        small = t[rest] < 1e-4
        # rows whose step became tiny leave the loop without acceptance
        done_small = rest[small]
        cont = rest[~small]
        t[cont] *= beta
        if done_small.size:
            searching[done_small] = False
            fb = done_small[f01[done_small] < 1.1 * f1[done_small]]
            if fb.size:
                Xf = X0[fb] - alpha[fb, None] * G[fb]
                if d > start:
                    Xf[:, start:d] = _soft(Xf[:, start:d], lam[fb])
                X1[fb] = Xf
    # rows that ran out of max_iters without break use the same fallback rule
    left = np.where(searching)[0]
    if left.size:
        fb = left[f01[left] < 1.1 * f1[left]]
        if fb.size:
            Xf = X0[fb] - alpha[fb, None] * G[fb]
            if d > start:
                Xf[:, start:d] = _soft(Xf[:, start:d], lam[fb])
            X1[fb] = Xf
    return X1, n_eval


def update_with_mask(Y, A, B, d, lam, P1, P2, family, nu, Ys, thres_disp, a, C,
                     alpha, beta, max_iters, tol, gene_active, cell_active, alpha_gene,
                     stats=None):
    """One alternating iteration (``causarray/gcate_opt.py:148-206``).

    ``nu`` is ``(1, p)``; ``A`` and ``B`` are modified in place and returned.
    """
    n, p = Y.shape
    nu_g = nu.reshape(-1)
    # ---- A-step ----
    g = grad(Y.T, B, A, family, nu.T, thres_disp)          # (n, k)
    g[:, :d] = 0.0
    rows = np.where(cell_active)[0]
    if rows.size:
        g[rows, d:] = _ball(g[rows, d:], 2 * C)
        X1, ne = line_search_rows(Y[rows], B, A[rows], g[rows], d, family, nu_g, Ys[rows],
                                  thres_disp, 0, 0.0, np.full(rows.size, alpha), beta, max_iters, tol,
                                  nu_per_row=False)
        A[rows] = X1
        if stats is not None:
            stats["cell_evals"] = stats.get("cell_evals", 0) + int(ne.sum())
    if P1 is not None:
        A[:, d:] -= P1 @ (P1.T @ A[:, d:])
    # ---- B-step ----
    g = grad(Y, A, B, family, nu, thres_disp)              # (p, k)
    if P1 is not None:
        g[:, :d] = 0.0
    elif P2 is not None:
        g[:, d - a:d] -= P2 @ (P2.T @ g[:, d - a:d])
        g[:, d:] = 0.0
        g[:, :d - a] = 0.0
    cols = np.where(gene_active)[0]
    if cols.size:
        if P2 is None:
            g[cols, d:] = _ball(g[cols, d:], 2 * C)
        else:
            g[cols, :d] = _ball(g[cols, :d], 2 * C)
        X1, ne = line_search_rows(np.ascontiguousarray(Y[:, cols].T), A, B[cols], g[cols], d, family,
                                  nu_g[cols], np.ascontiguousarray(Ys[:, cols].T), thres_disp, d - a,
                                  lam[cols], alpha_gene[cols], beta, max_iters, tol, nu_per_row=True)
        B[cols] = X1
        if stats is not None:
            stats["gene_evals"] = stats.get("gene_evals", 0) + int(ne.sum())
    if P2 is None:
        B[:, d:] = np.clip(B[:, d:], -10.0, 10.0)
    f = nll(Y, A, B, family, nu, Ys, thres_disp)
    return f, A, B


class EarlyStopping:
    """``causarray/utils.py:101-161`` (minimisation only)."""

    def __init__(self, warmup=25, patience=25, tolerance=0.0, rel_tol=1e-4, **kw):
        self.warmup, self.patience, self.tolerance, self.rel_tol = warmup, patience, tolerance, rel_tol
        self.step, self.best_step, self.best_metric = -1, -1, np.inf

    def __call__(self, metric):
        self.step += 1
        if self.step < self.warmup:
            return False
        thr = self.tolerance + self.rel_tol * abs(self.best_metric) if np.isfinite(self.best_metric) else 0.0
        if metric < self.best_metric - thr:
            self.best_metric, self.best_step = metric, self.step
            return False
        return self.step - self.best_step > self.patience


def log_h_terms(Yn, nu, family, thres_disp):
    """``Ys`` of ``gcate_opt.py:345-347``: ``-lgamma(Y+1) [+ lgamma(nu+Y) - lgamma(nu)]``."""
    Ys = -gammaln(Yn + 1.0)
    if family == "nb":
        Ys = Ys + np.where(nu > thres_disp, 0.0, gammaln(nu + Yn) - gammaln(nu))
    return Ys


def alter_min_loop(Y, X, r, A0, B0, family, disp, size_factor, stage, P2=None, lam=0.0, a=None,
                   kwargs_ls=None, kwargs_es=None, thres_disp=100.0):
    """The optimiser part of ``alter_min`` (``gcate_opt.py:285-458``) with
    explicit initial ``A0 (n, d+r)``, ``B0 (p, d+r)`` (the GLM/SVD initialisation
    ``:308-334`` is restated separately).  ``stage`` 1 uses ``P1 = qr(X)``,
    stage 2 the supplied thin-Q ``P2``.
    """
    import scipy.linalg
    n, p = Y.shape
    d = X.shape[1]
    ls = {"alpha": 0.1, "beta": 0.5, "max_iters": 20, "tol": 1e-4, "C": None, "tol_gene": 1e-4,
          "tol_cell": 1e-4, "recheck_interval": 10, "sparsity_threshold": 0.5, "sparsity_boost": 2.0,
          "warmup_iters": 0}
    ls.update(kwargs_ls or {})
    es_kw = {"max_iters": 50, "warmup": 0, "patience": 5, "tolerance": 0.0, "rel_tol": 2e-4}
    es_kw.update(kwargs_es or {})
    C = ls["C"] if ls["C"] is not None else (1e3 if family == "nb" else 1e5)
    nu = np.asarray(disp, dtype=np.float64).reshape(1, -1)
    sf = np.asarray(size_factor, dtype=np.float64).reshape(-1, 1)
    P1 = None
    if stage == 1:
        P1 = scipy.linalg.qr(X, mode="economic")[0].astype(np.float64)
    if a is None:
        a = d
    sparsity = (Y == 0).mean(axis=0).astype(np.float64)
    A = np.array(A0, dtype=np.float64)
    B = np.array(B0, dtype=np.float64)
    if P2 is not None:
        P2 = np.asarray(P2, dtype=np.float64)
        B[:, d - a:d] -= P2 @ (P2.T @ B[:, d - a:d])
    Yn = Y.astype(np.float64) / sf
    Ys = log_h_terms(Yn, nu, family, thres_disp)
    weights = float(lam) / (np.abs(B[:, d - a:d]) + 1e-6)
    alpha_gene = np.full(p, ls["alpha"], dtype=np.float64)
    sm = sparsity > ls["sparsity_threshold"]
    alpha_gene[sm] *= 1.0 + ls["sparsity_boost"] * sparsity[sm]
    gene_active = np.ones(p, dtype=bool)
    cell_active = np.ones(n, dtype=bool)
    f_pre = (nll(Yn, A, B, family, nu, Ys, thres_disp) + np.sum(np.abs(B[:, d - a:d]) * weights)) / p
    f = f_pre
    hist = [f_pre]
    es = EarlyStopping(**es_kw)
    tot_g = tot_c = 0
    t = 0
    stats = {}
    for t in range(es_kw["max_iters"]):
        B_prev = B.copy()
        U_prev = A[:, d:].copy()
        f, A, B = update_with_mask(Yn, A, B, d, weights, P1, P2, family, nu, Ys, thres_disp, a, C,
                                   ls["alpha"], ls["beta"], ls["max_iters"], ls["tol"],
                                   gene_active, cell_active, alpha_gene, stats=stats)
        tot_g += int(gene_active.sum())
        tot_c += int(cell_active.sum())
        if ls["recheck_interval"] > 0 and (t + 1) % ls["recheck_interval"] == 0:
            gene_active = np.ones(p, dtype=bool)
            cell_active = np.ones(n, dtype=bool)
        else:
            if ls["tol_gene"] > 0:
                gene_active = np.linalg.norm(B - B_prev, axis=1) > ls["tol_gene"]
            if ls["tol_cell"] > 0:
                cell_active = np.linalg.norm(A[:, d:] - U_prev, axis=1) > ls["tol_cell"]
        f = (f + np.sum(np.abs(B[:, d - a:d]) * weights)) / p
        hist.append(f)
        if (not np.isfinite(f)) or f > max(1e3 * abs(f_pre), 1e3):
            break
        elif es(f):
            break
        else:
            f_pre = f
    return {"n_iter": t, "func_val": f, "resid": f_pre - f, "hist": hist, "X_U": A, "B_Gamma": B,
            "U": A[:, d:], "total_gene_updates": tot_g, "total_cell_updates": tot_c, "stats": stats}
