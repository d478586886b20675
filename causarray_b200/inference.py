"""Inference tail: p-values, Benjamini-Hochberg, empirical-null adjustment, FDX control.

Interface of the reference's ``causarray/DR_inference.py`` (``bh_correction``
``:153-201``, ``fdx_control`` ``:102-150``, ``multiplier_bootstrap`` ``:8-31``,
``step_down`` ``:34-70``, ``augmentation`` ``:73-99``).  These operate on (p,)
vectors per treatment -- host work; the
statsmodels ``multipletests(method='fdr_bh')`` call is replaced by the
step-up below.
"""
import warnings

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


import concurrent.futures as _cf
import os as _os
def _cores_per_rank():
    """Host cores this process may use: all of them, or its share when torchrun runs one process per GPU."""
    try:
        ranks = max(1, int(_os.environ.get('LOCAL_WORLD_SIZE', _os.environ.get('WORLD_SIZE', '1'))))
    except ValueError:
        ranks = 1
    return max(1, (_os.cpu_count() or 1) // ranks)


_N_THREADS = max(1, min(16, _cores_per_rank()))
_POOL = _cf.ThreadPoolExecutor(max_workers=_N_THREADS, thread_name_prefix='causarray_inf')
_MAD_NORMAL = 0.6744897501960817   # scipy's scale="normal": Phi^{-1}(3/4)


def _host_lib():
    """The shared library's host-side helpers (no GPU needed); ``None`` when it has not been built."""
    try:
        from . import _lib as L
        return L.load(require_gpu=False), L
    except Exception:
        return None, None


def _bh_rows(p):
    """Row-wise Benjamini-Hochberg, NaN entries left out of each family: ``ca_host_bh_rows`` (one host
    thread per row) when the library is built, else the numpy form below (same arithmetic)."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    lib, L = _host_lib()
    if lib is None or p.ndim != 2:
        return _bh_rows_block(p)
    q = np.empty_like(p)
    L.check(lib.ca_host_bh_rows(L.dptr(p), p.shape[0], p.shape[1], L.dptr(q)))
    return q


def _median_mad_rows(t):
    """NaN-skipping per-row median and MAD / Phi^{-1}(3/4) (scipy's ``scale='normal'``)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    lib, L = _host_lib()
    if lib is None:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            med = np.nanmedian(t, axis=1)
            mad = np.nanmedian(np.abs(t - med[:, None]), axis=1)
        return med, mad / _MAD_NORMAL
    med, mad = np.empty(t.shape[0]), np.empty(t.shape[0])
    L.check(lib.ca_host_median_mad_rows(L.dptr(t), t.shape[0], t.shape[1], L.dptr(med), L.dptr(mad)))
    return med, mad / _MAD_NORMAL


def _bh_rows_block(p):
    """Row-wise Benjamini-Hochberg with NaN entries left out of each family (vectorised numpy)."""
    a, m = p.shape
    order = np.argsort(p, axis=1)                      # NaNs sort last
    ps = np.take_along_axis(p, order, axis=1)
    cnt = np.sum(~np.isnan(p), axis=1)
    rank = np.arange(1, m + 1)[None, :]
    with np.errstate(invalid='ignore'):
        q = ps * cnt[:, None] / rank
    q = np.where(rank <= cnt[:, None], q, np.inf)
    q = np.minimum(np.minimum.accumulate(q[:, ::-1], axis=1)[:, ::-1], 1.0)
    q = np.where(rank <= cnt[:, None], q, np.nan)
    out = np.empty_like(q)
    np.put_along_axis(out, order, q, axis=1)
    return out


def _tail_parallel(x, dfa):
    """Two-sided tail probabilities of ``x`` (any shape), Student-t with ``dfa`` or normal, with the
    flat array cut into one contiguous piece per host thread (scipy.special ufuncs release the GIL;
    the Student-t tail -- an incomplete beta -- costs ~0.2 us per value)."""
    import scipy.special as sc
    xf = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.empty_like(xf)
    if dfa is None:
        np.multiply(sc.ndtr(-np.abs(xf)), 2.0, out=out)
        return out.reshape(np.shape(x))
    df_f = np.ascontiguousarray(np.broadcast_to(dfa, np.shape(x)), dtype=np.float64).ravel()
    cuts = np.linspace(0, xf.size, _N_THREADS + 1).astype(int)

    def work(i):
        lo, hi = cuts[i], cuts[i + 1]
        if hi > lo:
            sc.stdtr(df_f[lo:hi], -np.abs(xf[lo:hi]), out=out[lo:hi])
            out[lo:hi] *= 2.0
    list(_POOL.map(work, range(_N_THREADS)))
    return out.reshape(np.shape(x))


def bh_correction_device(tvalues, df=None):
    """``bh_correction`` of every row of ``tvalues (a, p)`` -- one BH family per treatment, as the
    reference loops over treatments (``causarray/DR_learner.py:225-231``) -- on the device
    (``ca_inference_tail``, ``csrc/inference.cu``): Student-t / normal tails, per-row bitonic sorts for
    the BH step-up and the median / MAD of the empirical null.  Raises without a CUDA device."""
    from . import _lib as L
    lib = L.load()
    t = np.ascontiguousarray(tvalues, dtype=np.float64)
    if t.ndim != 2:
        raise ValueError('tvalues must be (a, p)')
    dfa = None if df is None else np.ascontiguousarray(np.broadcast_to(np.asarray(df, dtype=np.float64), t.shape))
    out = [np.empty_like(t) for _ in range(4)]
    L.check(lib.ca_inference_tail(L.dptr(t), L.dptr(dfa), t.shape[0], t.shape[1], *(L.dptr(o) for o in out), None))
    return tuple(out)


def bh_correction_multi(tvalues, df=None):
    """``bh_correction`` applied to every row of ``tvalues (a, p)`` (one BH family per treatment, as
    the reference loops over treatments).  The tail probabilities call ``scipy.special.stdtr`` /
    ``ndtr`` directly -- the functions behind ``scipy.stats.t.sf`` / ``norm.sf`` -- without the
    distribution-object overhead, fanned out over host threads."""
    t = np.asarray(tvalues, dtype=np.float64)
    a = t.shape[0]
    keep = ~np.isnan(t)
    dfa = None if df is None else np.asarray(df, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        pv = np.where(keep, _tail_parallel(t, dfa), np.nan)
        med, mad = _median_mad_rows(np.where(keep, t, np.nan))
        ok = mad > 0
        t_adj = (t - med[:, None]) / np.where(ok, mad, 1.0)[:, None]
        keep_adj = keep & ok[:, None]
        pva = np.where(keep_adj, _tail_parallel(t_adj, dfa), np.nan)
        qv = _bh_rows(pv)
        qva = _bh_rows(pva)
    return pv, qv, pva, qva


def multiplier_bootstrap(eta, B):
    """Gaussian multiplier bootstrap statistics ``z = G eta`` with ``G (B, n)`` standard normal
    (``causarray/DR_inference.py:8-31`` draws the rows of ``G`` one by one from the same global
    numpy stream, so a seeded call reproduces the reference's statistics) -- one GEMM instead of
    ``B`` weighted column sums."""
    G = np.random.normal(size=(int(B), eta.shape[0]))
    return G @ np.asarray(eta, dtype=np.float64)


def step_down(tvalues_init, z_init, alpha):
    """Step-down max-|z| procedure (``causarray/DR_inference.py:34-70``): hypotheses leave in order
    of decreasing ``|t|`` while the largest remaining ``|t|`` reaches the ``1 - alpha`` quantile (lower
    interpolation) of the bootstrap maxima over the REMAINING hypotheses.

    The removal order is known in advance (descending ``|t|``, first index on ties -- what the
    reference's repeated ``argmax`` does), so the maxima over the remaining set are a reversed
    running maximum along that order and all thresholds come from one quantile call; the reference
    rescans the ``(B, p)`` matrix once per discovery.  Returns ``(V, tvalues, z)`` with the removed
    entries zeroed, like the reference."""
    t = np.array(tvalues_init, dtype=float)
    z = np.array(z_init, dtype=float)
    p = t.shape[0]
    V = np.zeros(p)
    if p == 0 or z.shape[0] == 0:
        return V, t, z
    order = np.argsort(-np.abs(t), kind='stable')
    absz = np.abs(z[:, order])
    rem_max = np.maximum.accumulate(absz[:, ::-1], axis=1)[:, ::-1]   # max over positions >= k of the order
    thr = np.quantile(rem_max, 1 - alpha, axis=0, method='lower')
    at = np.abs(t[order])
    stop = (at < thr) | (thr == 0)
    k = int(np.argmax(stop)) if stop.any() else p      # hypotheses order[:k] are discoveries
    hit = order[:k]
    V[hit] = 1
    t[hit] = 0
    z[:, hit] = 0
    return V, t, z


def augmentation(V, tvalues, c):
    """Augmentation step of FDX control (``causarray/DR_inference.py:73-99``): with ``|V|``
    step-down discoveries, the ``floor(c |V| / (1 - c))`` largest remaining ``|t|`` are added."""
    if c > 0:
        num_add = int(np.floor(c * np.sum(V) / (1 - c)))
        if num_add >= 1:
            at = np.abs(np.asarray(tvalues)).ravel()
            top = np.sort(at)[::-1][:num_add]
            V[np.isin(at, top)] = 1      # every hypothesis tied with one of the top values joins, as in the reference
    return V


def fdx_control(tau_est, tvalues_init, eta_est, std_est, fdx, B, alpha, c, min_var=1e-8, min_diff=0.1):
    """FDX (FDP exceedance) control: ``P(FDP > c) < alpha`` by multiplier bootstrap, step-down and
    augmentation (``causarray/DR_inference.py:102-150``).  Hypotheses with ``std < min_var`` are not
    tested; they count as discoveries when ``|tau| > min_diff``."""
    if not fdx:
        return np.zeros(np.shape(tvalues_init)[0])
    std_est = np.asarray(std_est, dtype=float)
    testable = std_est >= min_var
    with np.errstate(divide='ignore', invalid='ignore'):
        z = multiplier_bootstrap(np.asarray(eta_est) / std_est[None, :], B)
    z[:, ~testable] = 0.
    t = np.where(testable, tvalues_init, 0.)
    V, t, z = step_down(t, z, alpha)
    V[~testable & (np.abs(tau_est) > min_diff)] = 1.
    return augmentation(V, t, c)
