"""GCATE alternating minimisation on the device -- the S1 seam of SURVEY.md section 8(b).

Same names and argument meaning as the reference's numba layer:

* ``nll`` / ``grad``        ``causarray/gcate_likelihood.py:46-142``
* ``update`` / ``update_with_mask``   ``causarray/gcate_opt.py:89-206``
  (``A`` and ``B`` are updated in place and returned)
* ``alter_min``             ``causarray/gcate_opt.py:209-458`` (same result dict)

The numerics run in ``libcausarray_b200.so`` (``csrc/rowpass.cu``: fused DMMA
linear predictor + likelihood epilogue + backtracking line search); there is
no CPU path.
"""
import pprint

import numpy as np
import scipy.linalg

from . import glm as _glm_mod
from .device import GcateDevice, GlmDevice, RowsView, device_dtype
from .svd import topr_svd_device, topr_svd_native, lowrank_product_svd
from ._timing import phase

type_f = np.float64


def _family_check(family):
    if family not in ('poisson', 'nb'):
        raise ValueError('Family not recognized')


def _problem(Y, A, B, family, nuisance, Tys, thres_disp, d, r):
    """Device problem for a normalised response ``Y`` (rows = len(A), cols = len(B))."""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    nu = np.broadcast_to(np.asarray(nuisance, dtype=np.float64).reshape(-1), (p,)) if np.size(nuisance) in (1, p) \
        else None
    if nu is None:
        raise ValueError('nuisance must have one entry per feature')
    dev = GcateDevice(n, p, d, r, family, thres_disp)
    Ys = None
    if Tys is not None and np.shape(Tys) == Y.shape:
        Ys = np.asarray(Tys, dtype=np.float64)
    dev.load_normalized(Y, Ys, np.ascontiguousarray(nu))
    if Tys is not None and Ys is None and np.any(np.asarray(Tys) != 0):
        raise ValueError('Tys must be zeros or have the shape of Y')
    dev._tys_zero = Tys is not None and Ys is None
    dev.set_factors(np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64))
    return dev


def nll(Y, A, B, family, nuisance=np.ones((1, 1)), Tys=np.zeros((1, 1)), thres_disp=10.):
    """Negative log-likelihood ``-sum(Y*theta - b + Tys)/len(Y)`` for ``Theta = A @ B.T`` (2-D ``Y``)."""
    _family_check(family)
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError('the device nll takes a 2-D response')
    k = np.shape(A)[1]
    Tys_full = Tys if np.shape(Tys) == Y.shape else np.zeros_like(Y) + np.asarray(Tys, dtype=np.float64)
    dev = _problem(Y, A, B, family, nuisance, Tys_full, thres_disp, k, 0)
    return dev.nll()


def grad(Y, A, B, family, nuisance=np.ones((1, 1)), thres_disp=10.):
    """Gradient of the negative log-likelihood with respect to ``B``: ``(B.shape[0], k)``."""
    _family_check(family)
    Y = np.asarray(Y, dtype=np.float64)
    k = np.shape(A)[1]
    dev = _problem(Y, A, B, family, nuisance, np.zeros_like(Y), thres_disp, k, 0)
    return dev.grad(False)


def update_with_mask(Y, A, B, d, lam, P1, P2, family, nuisance, Ys, thres_disp, a, C,
                     alpha, beta, max_iters, tol, gene_active, cell_active, alpha_gene):
    """One alternating iteration with per-gene / per-cell convergence masks.

    Drop-in for the reference kernel: host arrays in, ``A`` and ``B`` updated in
    place, returns ``(func_val, A, B)``.  Every call uploads ``Y``; the resident
    loop is ``alter_min``.
    """
    _family_check(family)
    n, p = np.shape(Y)
    k = A.shape[1]
    dev = _problem(Y, A, B, family, nuisance, Ys, thres_disp, d, k - d)
    lam_arr = None
    if lam is not None and np.ndim(lam) == 2:
        lam_arr = np.ascontiguousarray(lam, dtype=np.float64)
        if lam_arr.shape != (p, a):
            raise ValueError('lam must have shape (p, a)')
        if not np.any(lam_arr != 0):
            lam_arr = None
    dev.set_stage(P1=P1, P2=P2, lam=lam_arr, a=a)
    dev.set_masks(gene_active, cell_active, alpha_gene)
    f = dev.update(C, alpha, beta, max_iters, tol)
    A_new, B_new = dev.get_factors()
    A[...] = A_new
    B[...] = B_new
    dev.close()
    return f, A, B


def update(Y, A, B, d, lam, P1, P2, family, nuisance, Ys, thres_disp, a, C, alpha, beta, max_iters, tol):
    """``update_with_mask`` with all-active masks and a uniform initial step."""
    return update_with_mask(Y, A, B, d, lam, P1, P2, family, nuisance, Ys, thres_disp, a, C,
                            alpha, beta, max_iters, tol, None, None, None)


class BatchWorkspace:
    """Raw counts of one (cells x genes) problem resident on the device, shared by the
    GCATE optimiser, the per-gene GLM fits and the AIPW reductions.

    ``BatchWorkspace(Y, size_factor, disp, family, d, r)`` uploads and prepares in one go;
    ``BatchWorkspace.upload(Y, family, d, r)`` only uploads (and runs the input checks on the
    device) so that size factors / dispersions can be computed from the device copy before
    ``prepare``.
    """

    def __init__(self, Y, size_factor, disp, family, d, r, thres_disp=100., _defer=False, shard=None):
        if not hasattr(Y, "tocsr") and not isinstance(Y, RowsView):   # a scipy.sparse matrix is uploaded as CSR
            Y = np.asarray(Y)
            if not device_dtype(Y):          # int32 / float32 / ... are widened on the device
                Y = np.asarray(Y, dtype=np.float64)
        n, p = Y.shape
        self.n, self.p = n, p
        self.gcate = GcateDevice(n, p, d, r, family, thres_disp)
        # gene shard of a fit spread over ranks: (rank, world, p_total, nccl_unique_id); Y holds this
        # rank's genes only
        self.shard = shard
        if shard is not None:
            self.gcate.set_shard(shard[0], shard[1], shard[2], shard[3])
        self.glm = GlmDevice(n, p)
        self.checks = None
        self.prepared = False
        if _defer:
            self.checks = self.gcate.upload_counts(Y, check=True)
        else:
            self.gcate.upload_counts(Y, check=False)
            self.prepare(size_factor, disp)
        self.glm.share_counts(self.gcate)

    @classmethod
    def upload(cls, Y, family, d, r, thres_disp=100., shard=None):
        return cls(Y, None, None, family, d, r, thres_disp, _defer=True, shard=shard)

    def prepare(self, size_factor, disp):
        self.gcate.prepare(size_factor, disp)
        self.prepared = True

    def close(self):
        self.glm.close()
        self.gcate.close()


def alter_min(Y, r, X=None, P1=None, P2=None, A=None, B=None,
              kwargs_glm={}, kwargs_ls={}, kwargs_es={},
              lam=0., a=None, verbose=False, thres_disp=100., _ws=None):
    """Alternating minimisation of the latent-factor GLM (same contract as the reference).

    Returns the reference's result dict: ``n_iter, func_val, resid, hist,
    kwargs_glm, kwargs_ls, kwargs_es, total_gene_updates, total_cell_updates,
    X_U, B_Gamma, U``.
    """
    if not hasattr(Y, "tocsr") and not (isinstance(Y, RowsView) and _ws is not None):
        Y = np.asarray(Y)                    # (sparse counts / row views live on the device only: _ws holds them)
    n, p = Y.shape
    d = X.shape[1]
    assert d > 0
    kwargs_glm = {**{'family': 'poisson', 'disp_glm': np.ones(p), 'size_factor': np.ones(n)}, **kwargs_glm}
    kwargs_ls = {**{
        'alpha': 0.1, 'beta': 0.5, 'max_iters': 20, 'tol': 1e-4, 'C': None,
        'tol_gene': 1e-4, 'tol_cell': 1e-4, 'recheck_interval': 10,
        'sparsity_threshold': 0.5, 'sparsity_boost': 2.0, 'warmup_iters': 0,
    }, **kwargs_ls}
    kwargs_es = {**{'max_iters': 50, 'warmup': 0, 'patience': 5, 'tolerance': 0., 'rel_tol': 2e-4}, **kwargs_es}
    if kwargs_es['max_iters'] == 0:
        return A, B, {'n_iter': 0, 'func_val': 0., 'resid': 0., 'hist': [0.], 'kwargs_glm': kwargs_glm,
                      'kwargs_ls': kwargs_ls, 'kwargs_es': kwargs_es}
    family = kwargs_glm['family']
    _family_check(family)
    nuisance = np.asarray(kwargs_glm['disp_glm'], dtype=type_f).reshape(-1)
    size_factor = np.asarray(kwargs_glm['size_factor'], dtype=type_f).reshape(-1)
    if kwargs_ls['C'] is None:
        kwargs_ls['C'] = 1e3 if family == 'nb' else 1e5
    X = np.asarray(X, dtype=type_f)
    if P1 is True:
        P1 = scipy.linalg.qr(X, mode='economic')[0].astype(type_f)
    if a is None:
        a = d

    ws = _ws if _ws is not None else BatchWorkspace(Y, size_factor, nuisance, family, d, r, thres_disp)
    offset = np.log(size_factor)
    # ---- initialisation: GLM -> deviance residuals -> top-r SVD (gcate_opt.py:308-334) ----
    if A is None:
        if verbose:
            pprint.pprint('Estimating initial latent variables with GLMs...')
        with phase('gcate_init_glm'):
            Bx, fit, _, _, _ = _glm_mod.fit_glm_auto(Y, X, offset=offset, family=family, disp_glm=nuisance, maxiter=100,
                                                verbose=verbose, _glm=ws.glm, _lazy=True)
        with phase('gcate_init_svd'):
            out = None
            if ws.shard is None and r + 20 <= 64:
                out = topr_svd_native(ws.glm, n, p, r, zero_cols=fit.refit_mask)
            if out is None:     # gene shards (reductions over the process group) / very wide blocks / breakdown
                ptr = ws.glm.resid_device_ptr()
                out = topr_svd_device(ptr, n, p, r, zero_cols=fit.refit_mask, sharded=ws.shard is not None)
            u, s, vt = out
        if u.shape[1] < r:
            raise ValueError('The number of latent factors is larger than the rank of deviance residuals '
                             '({}). Try to decrease the value of r.'.format(u.shape[1]))
        u_proj = u - P1 @ (P1.T @ u) if P1 is not None else u
        A = np.c_[X, u_proj]
    else:
        assert A.shape[1] == d + r
    A = np.array(A, dtype=type_f)
    if B is None:
        if verbose:
            pprint.pprint('Estimating initial coefficients with GLMs...')
        with phase('gcate_init_glm'):
            B = _glm_mod.fit_glm_auto(Y, A, offset=offset, family=family, disp_glm=nuisance, maxiter=100, verbose=verbose,
                                 _glm=ws.glm, _lazy=True)[0]
        u, s, vt = lowrank_product_svd(A[:, d:], B[:, d:], sharded=ws.shard is not None)
        A[:, d:] = u * s[None, :] ** (1 / 2)
        B[:, d:] = vt.T * s[None, :] ** (1 / 2)
    B = np.array(B, dtype=type_f)
    assert ~np.any(np.isnan(A))
    assert ~np.any(np.isnan(B))

    g = ws.gcate
    g.set_factors(A, B)
    g.set_stage(P1=P1, P2=None if P2 is None else np.asarray(P2, dtype=type_f), lam=None, a=a)
    if verbose:
        pprint.pprint({'kwargs_glm': kwargs_glm, 'kwargs_ls': kwargs_ls, 'kwargs_es': kwargs_es}, compact=True)
        pprint.pprint(f'Fitting GCATE (step {2 if P1 is None else 1})...')
    with phase('gcate_altmin_stage%d' % (2 if P1 is None else 1)):
        res_c, hist = g.alter_min(kwargs_ls, kwargs_es, lam, a)
        A, B = g.get_factors()
    kwargs_glm['disp_glm'] = nuisance
    res = {'n_iter': res_c.n_iter, 'func_val': res_c.func_val, 'resid': res_c.resid, 'hist': list(hist),
           'kwargs_glm': kwargs_glm, 'kwargs_ls': kwargs_ls, 'kwargs_es': kwargs_es,
           'total_gene_updates': int(res_c.total_gene_updates), 'total_cell_updates': int(res_c.total_cell_updates),
           'line_search_evals': (int(res_c.cell_evals), int(res_c.gene_evals)), 'loop_seconds': res_c.loop_seconds,
           'line_search_evals_onehot': (int(res_c.gene_evals_onehot), int(res_c.onehot_treated_cells))}
    res['X_U'] = A
    res['B_Gamma'] = B
    res['U'] = A[:, d:]
    if _ws is None:
        ws.close()
    return res
