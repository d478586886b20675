"""Thin object wrappers over the C-ABI handles (host numpy arrays in / out)."""
import ctypes

import numpy as np

from . import _lib as L


# element types ca_gcate_upload_counts_typed widens on the device (include/causarray_b200.h)
_NARROW_DTYPES = {np.dtype(np.float32).str: 1, np.dtype(np.int32).str: 2, np.dtype(np.int64).str: 3,
                  np.dtype(np.uint16).str: 4, np.dtype(np.uint8).str: 5}


def device_dtype(Y):
    """True when ``Y`` is a dense array the device upload takes without a host cast to float64."""
    return isinstance(Y, np.ndarray) and (Y.dtype == np.float64 or Y.dtype.str in _NARROW_DTYPES)


class RowsView:
    """``base[idx]`` without the copy: the rows ``idx`` of a C-contiguous host count matrix, as the
    batch drivers select them (``causarray/gcate.py:525-560``).  The device upload gathers the rows
    itself (``ca_gcate_upload_counts_rows``); anything else that needs the numbers gets the ordinary
    fancy-indexed copy through ``__array__`` / ``materialize``."""

    def __init__(self, base, idx):
        self.base = base
        self.idx = np.ascontiguousarray(idx, dtype=np.int64)
        self.shape = (int(self.idx.size), int(base.shape[1]))
        self.dtype = base.dtype
        self.ndim = 2
        self.direct = (isinstance(base, np.ndarray) and base.ndim == 2 and base.strides[1] == base.itemsize
                       and base.strides[0] >= base.shape[1] * base.itemsize
                       and (base.dtype == np.float64 or base.dtype.str in _NARROW_DTYPES))
        self._dense = None

    def materialize(self):
        if self._dense is None:
            self._dense = self.base[self.idx]
        return self._dense

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __len__(self):
        return self.shape[0]


class GcateDevice:
    """Device-resident GCATE problem (``ca_gcate_t``): S1 seam of SURVEY.md section 8b."""

    def __init__(self, n, p, d, r, family, thres_disp=100.0):
        if family not in L.FAMILY:
            raise ValueError('Family not recognized')
        self.lib = L.load()
        self.n, self.p, self.d, self.r, self.k = int(n), int(p), int(d), int(r), int(d + r)
        self.family = family
        self.h = ctypes.c_void_p()
        L.check(self.lib.ca_gcate_create(ctypes.byref(self.h), self.n, self.p, self.d, self.r,
                                         L.FAMILY[family], float(thres_disp)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ca_gcate_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank, world, p_total, unique_id):
        """Gene sharding over ``world`` ranks (one process per GPU): this handle holds ``p`` of
        ``p_total`` genes.  ``unique_id`` is the 128-byte NCCL id from :func:`comm_unique_id` of one
        rank.  Collective; call before the counts are loaded."""
        uid = np.ascontiguousarray(np.frombuffer(bytes(unique_id), dtype=np.uint8))
        assert uid.size == 128
        L.check(self.lib.ca_gcate_set_shard(self.h, int(rank), int(world), int(p_total), L.u8ptr(uid)))

    def transport(self):
        """How the cell-step reductions of a sharded handle travel: 'none', 'nccl' or 'nvlink-peer'."""
        return ('none', 'nccl', 'nvlink-peer')[int(self.lib.ca_gcate_transport(self.h))]

    def load_counts(self, Y, size_factor=None, nu=None):
        Y = L.f64(Y)
        assert Y.shape == (self.n, self.p)
        sf = None if size_factor is None else L.f64(np.ravel(size_factor))
        nu = None if nu is None else L.f64(np.ravel(nu))
        L.check(self.lib.ca_gcate_load_counts(self.h, L.dptr(Y), L.dptr(sf), L.dptr(nu)))

    def upload_counts(self, Y, check=True):
        if isinstance(Y, RowsView) and Y.direct:
            # rows of a larger host matrix: gathered inside the staged upload, no host copy of the batch
            c = [ctypes.c_int64(0) for _ in range(3)]
            refs = [ctypes.byref(x) for x in c] if check else [None, None, None]
            assert Y.shape == (self.n, self.p)
            code = 0 if Y.base.dtype == np.float64 else _NARROW_DTYPES[Y.base.dtype.str]
            L.check(self.lib.ca_gcate_upload_counts_rows(
                self.h, ctypes.c_void_p(Y.base.ctypes.data), int(Y.base.strides[0] // Y.base.itemsize),
                Y.idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), code, *refs))
            return tuple(int(x.value) for x in c)
        if isinstance(Y, RowsView):
            Y = np.asarray(Y)
        return self._upload_counts(Y, check)

    def _upload_counts(self, Y, check=True):
        """H2D of the raw counts (dense array, or a scipy.sparse matrix uploaded as CSR and expanded
        on the device); returns (#non-integer entries, #non-finite, #all-zero genes)."""
        c = [ctypes.c_int64(0) for _ in range(3)]
        refs = [ctypes.byref(x) for x in c] if check else [None, None, None]
        if hasattr(Y, "tocsr"):
            csr = Y.tocsr()
            if not csr.has_canonical_format:
                csr = csr.copy()
                csr.sum_duplicates()
            assert csr.shape == (self.n, self.p)
            indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
            data = np.ascontiguousarray(csr.data, dtype=np.float64)
            L.check(self.lib.ca_gcate_upload_counts_csr(
                self.h, indptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), L.i32ptr(indices), L.dptr(data),
                int(data.size), *refs))
        elif isinstance(Y, np.ndarray) and Y.dtype.str in _NARROW_DTYPES and Y.flags.c_contiguous:
            # int32 / float32 / ... counts cross the bus as they are and are widened on the device
            assert Y.shape == (self.n, self.p)
            L.check(self.lib.ca_gcate_upload_counts_typed(self.h, ctypes.c_void_p(Y.ctypes.data),
                                                          _NARROW_DTYPES[Y.dtype.str], *refs))
        else:
            Y = L.f64(Y)
            assert Y.shape == (self.n, self.p)
            L.check(self.lib.ca_gcate_upload_counts(self.h, L.dptr(Y), *refs))
        return tuple(int(x.value) for x in c)

    def size_factors(self):
        """Median-of-ratios size factors (``comp_size_factor(method='geomeans')``) from the device copy."""
        log_sf = np.empty(self.n)
        L.check(self.lib.ca_gcate_size_factor_medians(self.h, L.dptr(log_sf)))
        return np.exp(log_sf - np.mean(log_sf))

    def prepare(self, size_factor=None, nu=None):
        sf = None if size_factor is None else L.f64(np.ravel(size_factor))
        nu = None if nu is None else L.f64(np.ravel(nu))
        L.check(self.lib.ca_gcate_prepare(self.h, L.dptr(sf), L.dptr(nu)))

    def load_normalized(self, Yn, Ys=None, nu=None):
        Yn = L.f64(Yn)
        Ys = None if Ys is None else L.f64(Ys)
        nu = None if nu is None else L.f64(np.ravel(nu))
        L.check(self.lib.ca_gcate_load_normalized(self.h, L.dptr(Yn), L.dptr(Ys), L.dptr(nu)))

    def set_factors(self, A=None, B=None):
        A = None if A is None else L.f64(A)
        B = None if B is None else L.f64(B)
        if A is not None:
            assert A.shape == (self.n, self.k)
        if B is not None:
            assert B.shape == (self.p, self.k)
        L.check(self.lib.ca_gcate_set_factors(self.h, L.dptr(A), L.dptr(B)))

    def get_factors(self):
        A = np.empty((self.n, self.k))
        B = np.empty((self.p, self.k))
        L.check(self.lib.ca_gcate_get_factors(self.h, L.dptr(A), L.dptr(B)))
        return A, B

    def set_stage(self, P1=None, P2=None, lam=None, a=0):
        P1 = None if P1 is None else L.f64(P1)
        P2 = None if P2 is None else L.f64(P2)
        lam = None if lam is None else L.f64(lam)
        if lam is not None:
            assert lam.shape == (self.p, a)
        L.check(self.lib.ca_gcate_set_stage(self.h, L.dptr(P1), 0 if P1 is None else P1.shape[1],
                                            L.dptr(P2), 0 if P2 is None else P2.shape[1], L.dptr(lam), int(a)))

    def set_masks(self, gene_active=None, cell_active=None, alpha_gene=None):
        ga = None if gene_active is None else np.ascontiguousarray(gene_active, dtype=np.uint8)
        ca = None if cell_active is None else np.ascontiguousarray(cell_active, dtype=np.uint8)
        ag = None if alpha_gene is None else L.f64(alpha_gene)
        L.check(self.lib.ca_gcate_set_masks(self.h, L.u8ptr(ga), L.u8ptr(ca), L.dptr(ag)))

    def get_masks(self):
        ga = np.empty(self.p, dtype=np.uint8)
        ca = np.empty(self.n, dtype=np.uint8)
        L.check(self.lib.ca_gcate_get_masks(self.h, L.u8ptr(ga), L.u8ptr(ca)))
        return ga.astype(bool), ca.astype(bool)

    def update(self, C, alpha, beta, max_iters, tol):
        f = ctypes.c_double(0.0)
        L.check(self.lib.ca_gcate_update(self.h, float(C), float(alpha), float(beta), int(max_iters), float(tol),
                                         ctypes.byref(f)))
        return f.value

    def nll(self):
        f = ctypes.c_double(0.0)
        L.check(self.lib.ca_gcate_nll(self.h, ctypes.byref(f)))
        return f.value

    def grad(self, wrt_A):
        out = np.empty((self.n if wrt_A else self.p, self.k))
        L.check(self.lib.ca_gcate_grad(self.h, 1 if wrt_A else 0, L.dptr(out)))
        return out

    def alter_min(self, kwargs_ls, kwargs_es, lam, a):
        prm = L.AltminParams(
            alpha=kwargs_ls['alpha'], beta=kwargs_ls['beta'], tol=kwargs_ls['tol'], C=kwargs_ls['C'],
            ls_max_iters=int(kwargs_ls['max_iters']), tol_gene=kwargs_ls['tol_gene'], tol_cell=kwargs_ls['tol_cell'],
            recheck_interval=int(kwargs_ls['recheck_interval']), sparsity_threshold=kwargs_ls['sparsity_threshold'],
            sparsity_boost=kwargs_ls['sparsity_boost'], warmup_iters=int(kwargs_ls['warmup_iters']),
            es_max_iters=int(kwargs_es['max_iters']), es_warmup=int(kwargs_es['warmup']),
            es_patience=int(kwargs_es['patience']), es_tolerance=kwargs_es['tolerance'],
            es_rel_tol=kwargs_es['rel_tol'], lam=float(lam), a=int(a))
        res = L.AltminResult()
        hist = np.zeros(int(kwargs_es['max_iters']) + 1)
        L.check(self.lib.ca_gcate_alter_min(self.h, ctypes.byref(prm), ctypes.byref(res), L.dptr(hist)))
        return res, hist[:res.hist_len].copy()


def comm_unique_id():
    """128-byte NCCL unique id (generate on one rank, broadcast to the others)."""
    out = np.zeros(128, dtype=np.uint8)
    L.check(L.load().ca_comm_unique_id(L.u8ptr(out)))
    return out.tobytes()


class GlmDevice:
    """Batched per-gene GLM fits (``ca_glm_t``): S2 seam."""

    def __init__(self, n, p):
        self.lib = L.load()
        self.n, self.p = int(n), int(p)
        self.h = ctypes.c_void_p()
        L.check(self.lib.ca_glm_create(ctypes.byref(self.h), self.n, self.p))
        self._keep = None

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ca_glm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_counts(self, Y):
        Y = L.f64(Y)
        assert Y.shape == (self.n, self.p)
        L.check(self.lib.ca_glm_load_counts(self.h, L.dptr(Y)))

    def share_counts(self, gcate):
        self._keep = gcate  # the GCATE handle owns the device copy
        L.check(self.lib.ca_glm_share_counts(self.h, gcate.h))

    def fit(self, family, X, offset=None, nu=None, thres_disp=100.0, maxiter=100, tol=1e-8, d_guard=None,
            ridge=None, gene_mask=None):
        """IRLS fit of all genes.  ``d_guard=None`` disables the second coefficient bound,
        ``d_guard=-1`` disables the guard altogether; ``ridge`` (kx,) / ``gene_mask`` (p,)
        select the penalised refit of a gene subset (see ``ca_glm_fit_ex``)."""
        if family not in L.FAMILY:
            raise ValueError('Family not recognized')
        X = L.f64(X)
        kx = X.shape[1]
        off = None if offset is None else L.f64(np.ravel(offset))
        nu = None if nu is None else L.f64(np.ravel(nu))
        rd = None if ridge is None else L.f64(np.ravel(ridge))
        gm = None if gene_mask is None else np.ascontiguousarray(gene_mask, dtype=np.uint8)
        B = np.zeros((self.p, kx))
        dev = np.zeros(self.p)
        n_iter = np.zeros(self.p, dtype=np.int32)
        status = np.zeros(self.p, dtype=np.int32)
        L.check(self.lib.ca_glm_fit_ex(self.h, L.FAMILY[family], L.dptr(X), kx, L.dptr(off), L.dptr(nu),
                                       float(thres_disp), int(maxiter), float(tol),
                                       int(kx if d_guard is None else d_guard), L.dptr(rd), L.u8ptr(gm),
                                       L.dptr(B), L.dptr(dev), L.i32ptr(n_iter), L.i32ptr(status)))
        return B, dev, n_iter, status

    def fit_treatment(self, family, X, B_cov, offset=None, nu=None, thres_disp=100.0, maxiter=50, tol=1e-8):
        """Treatment-effect fits against the covariate offset (``ca_glm_fit_treatment``): ``X (n, kx)``
        holds the covariates and a one-hot treatment block; the coefficients of the non-treatment
        columns are taken from ``B_cov (p, kx)`` (treatment entries = starting values) and kept."""
        X = L.f64(X)
        kx = X.shape[1]
        off = None if offset is None else L.f64(np.ravel(offset))
        nu = None if nu is None else L.f64(np.ravel(nu))
        B = np.array(B_cov, dtype=np.float64, order="C", copy=True)
        assert B.shape == (self.p, kx)
        dev = np.zeros(self.p)
        n_iter = np.zeros(self.p, dtype=np.int32)
        status = np.zeros(self.p, dtype=np.int32)
        L.check(self.lib.ca_glm_fit_treatment(self.h, L.FAMILY[family], L.dptr(X), kx, L.dptr(off), L.dptr(nu),
                                              float(thres_disp), int(maxiter), float(tol), L.dptr(B), L.dptr(dev),
                                              L.i32ptr(n_iter), L.i32ptr(status)))
        return B, dev, n_iter, status

    def refit_l1(self, family, X, l1, gene_mask, offset=None, nu=None, thres_disp=100.0, maxiter=100, tol=1e-8):
        """Lasso refit (``ca_glm_refit_l1``) of the genes flagged in ``gene_mask``, warm-started from the
        previous fit; ``l1`` (kx,) = n * alpha_k."""
        X = L.f64(X)
        kx = X.shape[1]
        off = None if offset is None else L.f64(np.ravel(offset))
        nu = None if nu is None else L.f64(np.ravel(nu))
        l1 = L.f64(np.ravel(l1))
        gm = np.ascontiguousarray(gene_mask, dtype=np.uint8)
        B = np.zeros((self.p, kx))
        dev = np.zeros(self.p)
        n_iter = np.zeros(self.p, dtype=np.int32)
        status = np.zeros(self.p, dtype=np.int32)
        L.check(self.lib.ca_glm_refit_l1(self.h, L.FAMILY[family], L.dptr(X), kx, L.dptr(off), L.dptr(nu),
                                         float(thres_disp), int(maxiter), float(tol), L.dptr(l1), L.u8ptr(gm),
                                         L.dptr(B), L.dptr(dev), L.i32ptr(n_iter), L.i32ptr(status)))
        return B, dev, n_iter, status

    def resid_device_ptr(self):
        ptr = ctypes.c_void_p()
        L.check(self.lib.ca_glm_resid_device(self.h, ctypes.byref(ptr)))
        return ptr.value

    def resid_device_ptrs(self):
        """Device pointers of the residual matrix of the last fit, cell-major (n, p) and gene-major
        (p, ld): ``(ptr, ptr_t, ld)``."""
        ptr, ptr_t, ld = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64(0)
        L.check(self.lib.ca_glm_resid_device2(self.h, ctypes.byref(ptr), ctypes.byref(ptr_t), ctypes.byref(ld)))
        return ptr.value, ptr_t.value, int(ld.value)

    def resid(self):
        out = np.empty((self.n, self.p))
        L.check(self.lib.ca_glm_resid(self.h, L.dptr(out)))
        return out

    def fitted(self):
        out = np.empty((self.n, self.p))
        L.check(self.lib.ca_glm_fitted(self.h, L.dptr(out)))
        return out

    def disp_moments(self, offset=None):
        off = None if offset is None else L.f64(np.ravel(offset))
        out = np.empty(self.p)
        L.check(self.lib.ca_glm_disp_moments(self.h, L.dptr(off), L.dptr(out)))
        return out


def _lfc_params(usevar, thres_min, thres_diff, eps_var, yhat_cap=1e5):
    if usevar not in ('pooled', 'unequal'):
        raise ValueError('usevar must be either "pooled" or "unequal"')
    return L.LfcParams(usevar=1 if usevar == 'unequal' else 0, thres_min=float(thres_min),
                       thres_diff=float(thres_diff), eps_var=float(eps_var), yhat_cap=float(yhat_cap))


def _lfc_outputs(a, p):
    outs = [np.empty((a, p)) for _ in range(5)]
    est = np.empty((a, p), dtype=np.uint8)
    return outs, est


def aipw_lfc(Y, W, Bcov, Btrt, offset, size_factor, A, pi, cell_mask=None, usevar='unequal', thres_min=1e-2,
             thres_diff=1e-2, eps_var=1e-4, glm=None):
    """Fused AIPW + LFC estimand (S3 seam).  Returns dict of (a, p) arrays."""
    lib = L.load()
    W, Bcov, Btrt, A, pi = L.f64(W), L.f64(Bcov), L.f64(Btrt), L.f64(A), L.f64(pi)
    n, kw = W.shape
    p, a = Btrt.shape
    off = None if offset is None else L.f64(np.ravel(offset))
    sf = L.f64(np.ravel(size_factor))
    cm = None if cell_mask is None else np.ascontiguousarray(cell_mask, dtype=np.uint8)
    prm = _lfc_params(usevar, thres_min, thres_diff, eps_var)
    outs, est = _lfc_outputs(a, p)
    tail = [L.u8ptr(cm), ctypes.byref(prm)] + [L.dptr(o) for o in outs] + [L.u8ptr(est)]
    if glm is not None:
        L.check(lib.ca_aipw_lfc_shared(glm.h, kw, a, L.dptr(W), L.dptr(Bcov), L.dptr(Btrt), L.dptr(off), L.dptr(sf),
                                       L.dptr(A), L.dptr(pi), *tail))
    else:
        Y = L.f64(Y)
        L.check(lib.ca_aipw_lfc(n, p, kw, a, L.dptr(Y), L.dptr(W), L.dptr(Bcov), L.dptr(Btrt), L.dptr(off),
                                L.dptr(sf), L.dptr(A), L.dptr(pi), *tail))
    return dict(mean0=outs[0], mean1=outs[1], tau=outs[2], var=outs[3], df=outs[4], estimable=est.astype(bool))


def aipw_lfc_yhat(Y, Y_hat, size_factor, A, pi, cell_mask=None, usevar='unequal', thres_min=1e-2, thres_diff=1e-2,
                  eps_var=1e-4):
    lib = L.load()
    Y, Y_hat, A, pi = L.f64(Y), L.f64(Y_hat), L.f64(A), L.f64(pi)
    n, p = Y.shape
    a = A.shape[1]
    assert Y_hat.shape == (n, p, a, 2)
    sf = L.f64(np.ravel(size_factor))
    cm = None if cell_mask is None else np.ascontiguousarray(cell_mask, dtype=np.uint8)
    prm = _lfc_params(usevar, thres_min, thres_diff, eps_var)
    outs, est = _lfc_outputs(a, p)
    L.check(lib.ca_aipw_lfc_yhat(n, p, a, L.dptr(Y), L.dptr(Y_hat), L.dptr(sf), L.dptr(A), L.dptr(pi), L.u8ptr(cm),
                                 ctypes.byref(prm), *[L.dptr(o) for o in outs], L.u8ptr(est)))
    return dict(mean0=outs[0], mean1=outs[1], tau=outs[2], var=outs[3], df=outs[4], estimable=est.astype(bool))
