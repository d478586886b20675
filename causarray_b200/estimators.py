"""Public GCATE estimators (host orchestration, device numerics).

Interface of the reference's ``causarray/gcate.py``: ``fit_gcate`` (``:54-133``),
``estimate`` (``:136-193``), ``estimate_r`` (``:196-325``) and
``fit_gcate_batch`` (``:328-609``) with the same arguments, defaults, result
dicts and error behaviour.  One ``BatchWorkspace`` keeps the counts of a fit
on the device across the initial GLMs, both optimiser stages and (through
``res_2['_ws']``) the LFC stage.
"""
import contextlib
import math
import time
import warnings

import numpy as np
import pandas as pd
import scipy.linalg
import scipy.sparse as _sp
from scipy.special import gammaln

from . import glm as _glm_mod
from .optim import alter_min, BatchWorkspace, type_f
from .device import RowsView
from .prep import comp_size_factor, _filter_params, subsample_ctrl_cells, subsample_pert_cells
from ._timing import phase


def _check_input(Y, X, family, disp_glm, disp_family, offset, c1, _ws=None, r=0, _shard=None, _size_factor=None,
                 **kwargs):
    """Input validation, size factors and dispersion (reference ``_check_input``).

    The O(n p) passes (integer check, empty-gene check, median-of-ratios size factors, the
    Poisson fit behind the dispersion) run on the device copy of ``Y`` held by the returned
    workspace.
    """
    if not (X.ndim == 2 and Y.ndim == 2):
        raise ValueError("Input must have ndim of 2. Y.ndim: {}, X.ndim: {}.".format(Y.ndim, X.ndim))
    if family not in ('poisson', 'nb'):
        raise ValueError('Family not recognized')
    n, p = Y.shape
    ws = _ws if _ws is not None else BatchWorkspace.upload(Y, family, X.shape[1], int(r),
                                                            kwargs.get('thres_disp', 100.), shard=_shard)
    if ws.checks is None:
        ws.checks = (0, 0, 0)
    if ws.shard is not None:      # every rank must take the same branch below
        from .parallel import allreduce_sum_np
        ws.checks = tuple(int(v) for v in allreduce_sum_np(np.asarray(ws.checks, dtype=np.float64)))
    n_nonint, n_nonfin, n_empty = ws.checks
    if n_nonint > 0 or n_nonfin > 0:
        warnings.warn("Y is not integer-valued. It will be rounded to the nearest integer.")
        Y = np.round(Y.toarray() if _sp.issparse(Y) else Y)
        ws.gcate.upload_counts(Y, check=False)
    if n_empty > 0:
        raise ValueError("Y contains non-expressed features.")
    if np.linalg.svd(X, compute_uv=False)[-1] < 1e-3:
        raise ValueError("The covariate matrix is near singular.")
    kwargs_glm = {'family': family}
    if offset is not None:
        if type(offset) == bool and offset is True:
            # (a gene shard cannot take the per-cell median over all genes: the caller supplies it)
            size_factor = ws.gcate.size_factors() if _size_factor is None else np.asarray(_size_factor, dtype=type_f)
            kwargs_glm['size_factor'] = size_factor
            offset = np.log(size_factor)
        else:
            # (the reference keeps size factors of one for an explicit / False offset)
            offset = None if offset is False else np.asarray(offset)
    if family == 'nb':
        if disp_family is None:
            disp_family = 'poisson'
        if disp_glm is None:
            disp_glm = _glm_mod.estimate_disp_auto(Y, X, offset=offset, disp_family=disp_family, maxiter=1000,
                                                   _glm=ws.glm)
    if disp_glm is not None:
        kwargs_glm['disp_glm'] = disp_glm
    kwargs_glm = {**{'family': 'gaussian', 'disp_glm': np.ones((1, p)), 'size_factor': np.ones((n, 1))},
                  **kwargs_glm}
    lam1 = 0.05 if c1 is None else c1
    if not ws.prepared:
        ws.prepare(np.ravel(kwargs_glm['size_factor']), np.ravel(kwargs_glm['disp_glm']))
    return Y, kwargs_glm, lam1, ws


def estimate(Y, X, r, a, lam1, kwargs_glm, kwargs_ls_1, kwargs_es_1, kwargs_ls_2, kwargs_es_2,
             A_init=None, _ws=None, **kwargs):
    """Two-stage alternating minimisation: stage 1 with ``P1 = qr([X | A])``, stage 2 with
    ``P2 = qr(Gamma-hat)`` and the adaptive L1 penalty ``lam1`` on the treatment block."""
    valid = _filter_params(alter_min, kwargs)
    valid.pop('A', None)
    valid.pop('_ws', None)
    n, p = Y.shape
    ws = _ws if _ws is not None else BatchWorkspace(
        Y, np.ravel(kwargs_glm['size_factor']), np.ravel(kwargs_glm['disp_glm']), kwargs_glm['family'],
        X.shape[1], int(r), valid.get('thres_disp', 100.))
    res_1 = alter_min(Y, r, X=X, P1=True, A=A_init, kwargs_glm=kwargs_glm, kwargs_ls=kwargs_ls_1,
                      kwargs_es=kwargs_es_1, _ws=ws, **valid)
    if ws.shard is not None:
        # thin Q of Gamma-hat with its rows (genes) split over ranks: Cholesky-QR through the all-reduced
        # r x r Gram matrix (P2 enters only through the projector Q Q^T, so the basis is immaterial)
        from .parallel import allreduce_sum_np
        Gm = res_1['B_Gamma'][:, -r:]
        Rc = np.linalg.cholesky(allreduce_sum_np(Gm.T @ Gm)).T
        Q = np.linalg.solve(Rc.T, Gm.T).T
    else:
        Q, _ = scipy.linalg.qr(res_1['B_Gamma'][:, -r:], mode='economic')
    if lam1 == 0.:
        res_2 = {'X_U': res_1['X_U'], 'B_Gamma': res_1['B_Gamma']}
    else:
        res_2 = alter_min(Y, r, X=X, P2=Q, A=res_1['X_U'].copy(), B=res_1['B_Gamma'].copy(), lam=lam1, a=a,
                          kwargs_glm=res_1['kwargs_glm'], kwargs_ls=kwargs_ls_2, kwargs_es=kwargs_es_2,
                          _ws=ws, **valid)
    res_2['_ws'] = ws
    return res_1, res_2


def fit_gcate(Y, X, A, r, family='nb', disp_glm=None, disp_family=None, offset=True,
              kwargs_ls_1={}, kwargs_ls_2={}, kwargs_es_1={}, kwargs_es_2={},
              c1=None, backend: str = "auto", A_init=None, _ws=None, gene_shard=False, **kwargs):
    """Fit GCATE: returns ``(res_1, res_2)``; ``res_2['X_U']`` is ``[X | A | U]`` and
    ``res_2['B_Gamma']`` the gene-level coefficients.  ``res_2['_ws']`` holds the device
    workspace of the fit (counts stay resident for a following :func:`LFC`); close it with
    ``res_2.pop('_ws').close()`` or let it be garbage collected."""
    X = np.hstack((np.asarray(X), np.asarray(A)))
    a = np.asarray(A).shape[1]
    shard_kw = {}
    if gene_shard:
        # One fit with its GENES split over the ranks of the default process group (one process per
        # GPU; every rank makes the same call with the full inputs and keeps its block of columns).
        # ``res['B_Gamma']`` then holds this rank's rows, ``res['gene_slice']`` says which.
        from .parallel import dist_info, gene_bounds, broadcast_bytes
        from .device import comm_unique_id
        rank, world = dist_info()
        if world > 1:
            Yf = _dense_counts(Y)
            p_total = Yf.shape[1]
            lo, hi = gene_bounds(p_total, world)[rank]
            uid = broadcast_bytes(comm_unique_id() if rank == 0 else None)
            shard_kw['_shard'] = (rank, world, p_total, uid)
            if offset is True:
                shard_kw['_size_factor'] = comp_size_factor(Yf)     # per-cell median over ALL genes
            Y = np.ascontiguousarray(Yf[:, lo:hi])
            if disp_glm is not None:
                disp_glm = np.ravel(disp_glm)[lo:hi]
            if A_init is not None:
                A_init = np.asarray(A_init)
    ctx = _glm_mod._backend_override(backend) if backend != "auto" else contextlib.nullcontext()
    with ctx:
        with phase('gcate_upload_checks_sizefactors'):
            Y, kwargs_glm, lam1, ws = _check_input(Y if (_sp.issparse(Y) or isinstance(Y, RowsView)) else np.asarray(Y),
                                                   X, family, disp_glm,
                                                   disp_family, offset, c1, _ws=_ws, r=int(r), **shard_kw, **kwargs)
        res_1, res_2 = estimate(Y, X, int(r), a, lam1, kwargs_glm, kwargs_ls_1, kwargs_es_1, kwargs_ls_2,
                                kwargs_es_2, A_init=A_init, _ws=ws, **kwargs)
    if shard_kw.get('_shard') is not None:
        rank, world, p_total, _ = shard_kw['_shard']
        from .parallel import gene_bounds
        res_1['gene_slice'] = res_2['gene_slice'] = gene_bounds(p_total, world)[rank]
    return res_1, res_2


def log_h(y, family, nuisance):
    if family == 'nb':
        return gammaln(y + nuisance) - gammaln(nuisance) - gammaln(y + 1)
    elif family == 'poisson':
        return -gammaln(y + 1)


def estimate_r(Y, X, A, r_max, c=1., family='nb', disp_glm=None, disp_family='poisson', offset=True,
               max_cells=None, random_state=0, kwargs_ls_1={}, kwargs_ls_2={}, kwargs_es_1={}, kwargs_es_2={},
               **kwargs):
    """JIC sweep over the number of latent factors (stage-1 fits only); returns the
    reference's DataFrame ``r, deviance, nu, JIC``."""
    from .device import GcateDevice
    A_np = np.asarray(A) if not isinstance(A, pd.DataFrame) else A.values
    n_total = np.asarray(Y).shape[0]
    if max_cells is not None and n_total > max_cells:
        rng = np.random.default_rng(random_state)
        ctrl_mask = A_np.sum(axis=1) == 0
        ctrl_idx, pert_idx = np.where(ctrl_mask)[0], np.where(~ctrl_mask)[0]
        n_pert_floor = min(len(pert_idx), max(1, max_cells // 4))
        n_ctrl_use = min(len(ctrl_idx), max_cells - n_pert_floor)
        n_pert_use = min(len(pert_idx), max_cells - n_ctrl_use)
        sel_c = rng.choice(ctrl_idx, n_ctrl_use, replace=False) if n_ctrl_use > 0 else np.array([], dtype=int)
        sel_p = rng.choice(pert_idx, n_pert_use, replace=False) if n_pert_use > 0 else np.array([], dtype=int)
        sel = np.sort(np.concatenate([sel_c, sel_p]))
        Y = Y.iloc[sel] if isinstance(Y, pd.DataFrame) else np.asarray(Y)[sel]
        X = np.asarray(X)[sel]
        A_np = A_np[sel]
        if offset is not True and offset is not False and offset is not None:
            offset = np.asarray(offset)[sel]
    a, d = A_np.shape[1], np.asarray(X).shape[1]
    X = np.hstack((np.asarray(X), A_np))
    Y = Y.values if isinstance(Y, pd.DataFrame) else np.asarray(Y)
    n, p = Y.shape
    Y, kwargs_glm, _, ws = _check_input(Y, X, family, disp_glm, disp_family, offset, None, r=int(np.max(r_max)), **kwargs)
    family = kwargs_glm['family']
    nuisance = np.asarray(kwargs_glm['disp_glm'], dtype=type_f).reshape(1, -1)
    size_factor = np.asarray(kwargs_glm['size_factor'], dtype=type_f).reshape(-1, 1)
    if np.isscalar(r_max):
        r_list = np.arange(1, int(r_max) + 1)
    else:
        r_list = np.array(r_max, dtype=int)
    r_top = int(np.max(r_list))

    from .svd import topr_svd_device
    B0, fit, _, _, _ = _glm_mod.fit_glm(Y, X, offset=np.log(size_factor[:, 0]), family=family,
                                        disp_glm=nuisance[0], maxiter=100, _glm=ws.glm, _lazy=True)
    u, s, vt = topr_svd_device(ws.glm.resid_device_ptr(), n, p, r_top, zero_cols=fit.refit_mask)
    if u.shape[1] < r_top:
        raise ValueError('The number of latent factors is larger than the rank of deviance residuals '
                         '({}). Try to decrease the value of r.'.format(u.shape[1]))
    Q = scipy.linalg.qr(X, mode='economic')[0].astype(type_f)
    A1 = np.c_[X, u - Q @ (Q.T @ u)]
    logh_sum = np.sum(log_h(Y, family, nuisance))

    def _nll_raw(Amat, Bmat):
        # the reference passes size_factor positionally as Tys here (gcate.py:306,318): the
        # constant term is sum_i p * size_factor_i and Y is NOT normalised
        k = Amat.shape[1]
        dev = GcateDevice(n, p, k, 0, family, 100.)
        dev.load_normalized(Y, np.zeros_like(Y), nuisance[0])
        dev.set_factors(Amat, Bmat)
        f = dev.nll()
        dev.close()
        return f - np.sum(size_factor) * p / n

    rows = []
    ll = 2 * (_nll_raw(X, B0) / p - logh_sum / (n * p))
    nu = (d + a) * np.maximum(n, p) * np.log(n * p / np.maximum(n, p)) / (n * p)
    rows.append([0, ll, nu, ll + c * nu])
    for r in r_list[r_list > 0][::-1]:
        # The reference passes ``A=A1[:, :d+a+r]`` here, but ``estimate`` drops a stray ``A=`` keyword
        # (causarray/gcate.py:176-178), so every r starts from its own GLM + SVD initialisation; the
        # same happens here (no ``A_init``) -- a warm start would lower the deviances of r < r_max.
        _, res_2 = estimate(Y, X, int(r), a, 0, kwargs_glm, kwargs_ls_1, kwargs_es_1, kwargs_ls_2, kwargs_es_2,
                            **{k_: v for k_, v in kwargs.items() if k_ != 'A'})
        res_2.pop('_ws').close()
        A1, A2 = res_2['X_U'], res_2['B_Gamma']
        ll = 2 * (_nll_raw(A1, A2) / p - logh_sum / (n * p))
        nu = (d + a + r) * np.maximum(n, p) * np.log(n * p / np.maximum(n, p)) / (n * p)
        rows.append([r, ll, nu, ll + c * nu])
    ws.close()
    return pd.DataFrame(rows, columns=['r', 'deviance', 'nu', 'JIC']).sort_values(by='r')


def _dense_counts(Y):
    if _sp.issparse(Y):
        gb = Y.shape[0] * Y.shape[1] * 8 / 1e9
        if gb > 4.0:
            warnings.warn(f"Materialising sparse Y ({Y.shape[0]}x{Y.shape[1]}) into dense float64 requires "
                          f"~{gb:.1f} GB of memory.", ResourceWarning, stacklevel=3)
        return np.asarray(Y.toarray(), dtype=np.float64)
    if isinstance(Y, pd.DataFrame):
        return Y.values
    return np.asarray(Y)


def fit_gcate_batch(Y, X, A, r, batch_size=10, n_batches=None, max_cells=2000, n_ctrl=2000, family='nb',
                    disp_glm=None, disp_family=None, offset=True, warm_start_U=False, skip_batches=None,
                    random_state=0, verbose=False, **kwargs):
    """GCATE on chunks of perturbations against one fixed control subsample.

    Returns the reference's list of per-batch dicts (``batch_i, pert_names, ctrl_idx,
    pert_idx, cell_idx, res_1, res_2, disp_glm, t_batch[, skipped]``).
    """
    out = []
    for br in iter_gcate_batches(Y, X, A, r, batch_size=batch_size, n_batches=n_batches, max_cells=max_cells,
                                 n_ctrl=n_ctrl, family=family, disp_glm=disp_glm, disp_family=disp_family,
                                 offset=offset, warm_start_U=warm_start_U, skip_batches=skip_batches,
                                 random_state=random_state, verbose=verbose, **kwargs):
        if br.get('res_2') is not None and '_ws' in br['res_2']:
            br['res_2'].pop('_ws').close()
        out.append(br)
    return out


def iter_gcate_batches(Y, X, A, r, batch_size=10, n_batches=None, max_cells=2000, n_ctrl=2000, family='nb',
                       disp_glm=None, disp_family=None, offset=True, warm_start_U=False, skip_batches=None,
                       random_state=0, verbose=False, **kwargs):
    """Generator behind :func:`fit_gcate_batch`: yields one batch dict at a time with the
    device workspace still attached as ``res_2['_ws']`` (the caller closes it)."""
    # A sparse count matrix stays sparse: only the rows of one batch (and of the control subsample)
    # are densified at a time, so host memory is bounded by a batch (reference: _maybe_densify on
    # the batch slice, causarray/gcate.py:451-470)
    sparse_in = _sp.issparse(Y)
    Y_np = Y.tocsr() if sparse_in else _dense_counts(Y)

    def rows(idx, dense=True):
        if not sparse_in:
            # dense=False: a view (the device upload gathers the rows itself, see RowsView)
            return Y_np[idx] if dense else RowsView(Y_np, idx)
        return np.asarray(Y_np[idx].toarray(), dtype=np.float64) if dense else Y_np[idx]
    X_np = np.asarray(X)
    if isinstance(A, pd.DataFrame):
        names = list(A.columns)
        A_np = A.values.astype(float)
    else:
        A_np = np.asarray(A, dtype=float)
        names = list(range(A_np.shape[1]))
    n, p = Y_np.shape
    a_total = A_np.shape[1]
    multi = int(np.sum(A_np.sum(axis=1) > 1))
    if multi > 0:
        warnings.warn(f"fit_gcate_batch detected {multi} cell(s) with more than one active perturbation. "
                      "Within each batch, off-chunk perturbations are treated as controls, which biases the "
                      "per-batch null for combinatorial designs.  For pure combinatorial screens pass "
                      "``n_batches=1`` (no chunking) or exclude multi-pert rows before calling.",
                      RuntimeWarning, stacklevel=2)
    ctrl_sel = subsample_ctrl_cells(np.where(A_np.sum(axis=1) == 0)[0], n_ctrl=n_ctrl, random_state=random_state)
    if family == 'nb' and disp_glm is None:
        if offset is True:
            off_c = np.log(comp_size_factor(rows(ctrl_sel)))
        elif offset is not None and offset is not False:
            off_c = np.asarray(offset)[ctrl_sel]
        else:
            off_c = None
        disp_glm = _glm_mod.estimate_disp_auto(rows(ctrl_sel), X_np[ctrl_sel], offset=off_c)
    if n_batches is None:
        n_batches = math.ceil(a_total / batch_size)
    n_batches = max(1, min(n_batches, a_total))
    chunks = [list(c) for c in np.array_split(range(a_total), n_batches) if len(c) > 0]
    skip = set(skip_batches) if skip_batches is not None else set()
    U_prev = None
    t_total = 0.0
    for batch_i, cols in enumerate(chunks):
        pert_names = [names[j] for j in cols]
        if batch_i in skip:
            yield {'batch_i': batch_i, 'pert_names': pert_names, 'ctrl_idx': ctrl_sel, 'pert_idx': None,
                   'cell_idx': None, 'res_1': None, 'res_2': None, 'disp_glm': disp_glm, 't_batch': 0.0,
                   'skipped': True}
            continue
        t0 = time.perf_counter()
        pert_idx = np.where(A_np[:, cols].sum(axis=1) > 0)[0]
        pert_sel = pert_idx if max_cells is None else subsample_pert_cells(
            pert_idx, max_cells=max_cells, random_state=random_state + batch_i)
        cell_idx = np.sort(np.concatenate([ctrl_sel, pert_sel]))
        # (a sparse batch slice goes to the device as CSR; nothing on the host needs its dense form)
        Y_b, X_b = rows(cell_idx, dense=False), X_np[cell_idx]
        A_b = A_np[cell_idx][:, cols]
        kw = dict(kwargs)
        kw['disp_glm'] = disp_glm
        if 'disp_family' not in kw:
            kw['disp_family'] = 'poisson' if disp_glm is not None else disp_family
        if warm_start_U and U_prev is not None:
            loc = np.searchsorted(cell_idx, ctrl_sel)
            d_b = X_b.shape[1] + len(cols)
            init = np.zeros((len(cell_idx), d_b + r), dtype=np.float32)
            init[:, :d_b] = np.c_[X_b, A_b]
            init[loc, d_b:] = U_prev
            kw['A_init'] = init
        res_1, res_2 = fit_gcate(Y_b, X_b, A_b, r, family=family, offset=offset, **kw)
        if warm_start_U:
            d_b = X_b.shape[1] + len(cols)
            loc = np.searchsorted(cell_idx, ctrl_sel)
            U_prev = res_2['X_U'][loc, d_b:].copy()
        t_batch = time.perf_counter() - t0
        t_total += t_batch
        if verbose:
            t_avg = t_total / (batch_i + 1)
            print(f'Batch {batch_i + 1}/{n_batches}: {len(pert_names)} perts, {len(cell_idx)} cells, '
                  f'{t_batch:.2f}s (avg {t_avg:.2f}s/batch)')
        yield {'batch_i': batch_i, 'pert_names': pert_names, 'ctrl_idx': ctrl_sel, 'pert_idx': pert_sel,
               'cell_idx': cell_idx, 'res_1': res_1, 'res_2': res_2, 'disp_glm': disp_glm, 't_batch': t_batch}
