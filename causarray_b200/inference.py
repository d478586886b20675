"""Inference tail: p-values, Benjamini-Hochberg, empirical-null adjustment, FDX control.

Interface of the reference's ``causarray/DR_inference.py`` (``bh_correction``
``:153-201``, ``fdx_control`` ``:102-150``, ``multiplier_bootstrap`` ``:8-31``,
``step_down`` ``:34-70``, ``augmentation`` ``:73-99``).  These operate on (p,)
vectors per treatment -- host work; the
statsmodels ``multipletests(method='fdr_bh')`` call is replaced by the
step-up below.
"""
import numpy as np
import scipy.stats


def _bh(p):
    m = p.size
    order = np.argsort(p)
    q = p[order] * m / np.arange(1, m + 1)
    q = np.minimum(np.minimum.accumulate(q[::-1])[::-1], 1.0)
    out = np.empty_like(q)
    out[order] = q
    return out


def bh_correction(tvalues_init, df=None):
    """Two-sided p-values (t with ``df`` or normal), BH q-values, and both again after the
    median / MAD empirical-null standardisation.  NaN statistics propagate as NaN."""
    t = np.asarray(tvalues_init, dtype=np.float64)
    keep = ~np.isnan(t)
    pvals, qvals, pvals_adj, qvals_adj = (np.full(t.shape, np.nan) for _ in range(4))
    if keep.sum() > 0:
        def two_sided(x):
            if df is not None:
                return scipy.stats.t.sf(np.abs(x), df=np.asarray(df)[keep]) * 2
            return scipy.stats.norm.sf(np.abs(x)) * 2
        pvals[keep] = two_sided(t[keep])
        qvals[keep] = _bh(pvals[keep])
        med = np.nanmedian(t[keep])
        mad = scipy.stats.median_abs_deviation(t[keep], scale="normal", nan_policy='omit')
        if mad > 0:
            t_adj = (t - med) / mad
            pvals_adj[keep] = two_sided(t_adj[keep])
            qvals_adj[keep] = _bh(pvals_adj[keep])
    return pvals, qvals, pvals_adj, qvals_adj


def multiplier_bootstrap(eta, B):
    """Gaussian multiplier bootstrap statistics ``(B, p)`` of the scaled influence values."""
    n, p = eta.shape
    G = np.random.normal(size=(int(B), n))
    return G @ eta


def step_down(tvalues_init, z_init, alpha):
    p = z_init.shape[1]
    V = np.zeros(p,)
    z = z_init.copy()
    t = tvalues_init.copy()
    while True:
        i = np.argmax(np.abs(t))
        t_max = np.abs(t[i])
        quan = np.quantile(np.max(np.abs(z), axis=1), 1 - alpha, method='lower')
        if t_max < quan or quan == 0:
            break
        t[i] = 0
        z[:, i] = 0
        V[i] = 1
    return V, t, z


def augmentation(V, tvalues, c):
    if c > 0:
        num_add = int(np.floor(c * np.sum(V) / (1 - c)))
        if num_add >= 1:
            srt = np.sort(np.abs(tvalues), axis=None)[::-1]
            for i in np.arange(num_add):
                V[np.abs(tvalues) == srt[i]] = 1
    return V


def fdx_control(tau_est, tvalues_init, eta_est, std_est, fdx, B, alpha, c, min_var=1e-8, min_diff=0.1):
    """FDX (FDP exceedance) control by multiplier bootstrap + step-down + augmentation."""
    if not fdx:
        return np.zeros(tvalues_init.shape[0])
    ok = std_est >= min_var
    z = multiplier_bootstrap(eta_est / std_est[None, :], B)
    z[:, ~ok] = 0.
    t = tvalues_init.copy()
    t[~ok] = 0.
    V, t, z = step_down(t, z, alpha)
    V[(~ok) & (np.abs(tau_est) > min_diff)] = 1.
    return augmentation(V, t, c)
