"""CPU restatement of the inference tail (TEST INFRASTRUCTURE).

Follows ``causarray/DR_inference.py:153-201`` (``bh_correction``) and the
Benjamini-Hochberg step-up of ``statsmodels.stats.multitest.multipletests(
method='fdr_bh')`` that it calls (``:186,199``): sort p, q_(i) = p_(i) m / i,
running minimum from the largest rank down, clip at 1, un-sort.
"""
import numpy as np
import scipy.stats


def fdr_bh(p):
    p = np.asarray(p, dtype=np.float64)
    m = p.size
    if m == 0:
        return p.copy()
    order = np.argsort(p)
    ps = p[order]
    q = ps * m / np.arange(1, m + 1)
    q = np.minimum.accumulate(q[::-1])[::-1]
    q = np.minimum(q, 1.0)
    out = np.empty_like(q)
    out[order] = q
    return out


def multipletests(pvals, alpha=0.05, method="fdr_bh", **kw):
    assert method == "fdr_bh"
    q = fdr_bh(pvals)
    return q <= alpha, q, None, None


def fdrcorrection(pvals, alpha=0.05, **kw):
    q = fdr_bh(pvals)
    return q <= alpha, q


def bh_correction(t, df=None):
    """``causarray/DR_inference.py:153-201``."""
    t = np.asarray(t, dtype=np.float64)
    idx = ~np.isnan(t)
    pv = np.full(t.shape, np.nan)
    qv = np.full(t.shape, np.nan)
    pa = np.full(t.shape, np.nan)
    qa = np.full(t.shape, np.nan)
    if idx.sum() > 0:
        sf = (lambda x: scipy.stats.t.sf(x, df=df[idx])) if df is not None else scipy.stats.norm.sf
        pv[idx] = sf(np.abs(t[idx])) * 2
        qv[idx] = fdr_bh(pv[idx])
        med = np.nanmedian(t[idx])
        mad = scipy.stats.median_abs_deviation(t[idx], scale="normal", nan_policy="omit")
        if mad > 0:
            ta = (t - med) / mad
            pa[idx] = sf(np.abs(ta[idx])) * 2
            qa[idx] = fdr_bh(pa[idx])
    return pv, qv, pa, qa
