"""Top-r SVD of the deviance-residual matrix for the GCATE initialisation.

Replaces ``scipy.sparse.linalg.svds(resid, k=r)`` (``causarray/gcate_opt.py:312``;
ARPACK, unseeded) by a block subspace iteration (CholeskyQR2 orthogonalisation) with Rayleigh-Ritz extraction
run on the device-resident residual matrix.  This is a one-off, plain
tall-skinny GEMM + QR workload, so it uses torch (cuBLAS / cuSOLVER) as
plumbing on the library's own device buffer (zero copy through
``__cuda_array_interface__``); it is not one of the hand-written hot kernels.

Singular vectors are defined up to sign; the convention here is "largest
|entry| of each left vector is positive".  The GCATE updates are sign
equivariant, so LFC results do not depend on it.
"""
import numpy as np


class _DevArray:
    """Expose a raw device pointer to torch via the CUDA array interface."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2, "strides": None}


def fix_signs(u, vt=None):
    for i in range(u.shape[1]):
        j = np.argmax(np.abs(u[:, i]))
        if u[j, i] < 0:
            u[:, i] *= -1
            if vt is not None:
                vt[i] *= -1
    return u, vt


def topr_svd_native(glm, n, p, r, zero_cols=None, oversample=20, max_iter=400, tol=1e-11, seed=0):
    """Top-``r`` singular triplets of the residual matrix of ``glm``'s last fit through the library's own
    subspace iteration (``ca_svd_topr``, ``csrc/svd.cu``: hand-written FP64 DMMA products, Cholesky-QR2,
    Rayleigh-Ritz; no torch / cuBLAS / cuSOLVER).  Returns ``u (n, r)``, ``s (r,)`` ascending, ``vt (r, p)``
    like ``svds``, or ``None`` when the block iteration broke down numerically (the caller falls back)."""
    import ctypes
    from . import _lib as L
    lib = L.load()
    ptr, ptr_t, ld = glm.resid_device_ptrs()
    rr = int(min(r, min(n, p)))
    U, S, Vt = np.empty((n, rr)), np.empty(rr), np.empty((rr, p))
    zc = None if zero_cols is None or not np.any(zero_cols) else np.ascontiguousarray(zero_cols, dtype=np.uint8)
    it = np.zeros(1, dtype=np.int32)
    rc = lib.ca_svd_topr(ctypes.c_void_p(ptr), ctypes.c_void_p(ptr_t), ld, n, p, rr, int(oversample), int(max_iter),
                         float(tol), int(seed), L.u8ptr(zc), L.dptr(U), L.dptr(S), L.dptr(Vt), L.i32ptr(it))
    if rc == -3:
        return None
    L.check(rc)
    u, s, vt = U[:, ::-1].copy(), S[::-1].copy(), Vt[::-1].copy()    # ascending, as svds returns them
    u, vt = fix_signs(u, vt)
    return u, s, vt


def topr_svd_device(ptr, n, p, r, zero_cols=None, oversample=20, max_iter=400, tol=1e-11, seed=0, sharded=False):
    """Top-``r`` singular triplets of the (n, p) float64 matrix at device pointer ``ptr``.

    Returns host arrays ``u (n, r)``, ``s (r,)`` (ascending, as ``svds``), ``vt (r, p)``.
    ``sharded``: the matrix is this rank's block of COLUMNS (genes) of the full matrix; sums over
    genes -- the (n, b) products ``R V``, the (b, b) Gram matrices of the Cholesky-QR and the Ritz
    residual -- are all-reduced over the process group, ``u`` / ``s`` come back replicated and ``vt``
    holds this rank's columns.
    """
    from .parallel import allreduce_sum, dist_info
    shard = sharded and dist_info()[1] > 1
    red = allreduce_sum if shard else (lambda t: t)
    import torch
    if not torch.cuda.is_available():
        from ._lib import CausarrayCudaError
        raise CausarrayCudaError("no CUDA device for the initial SVD")
    R = torch.as_tensor(_DevArray(ptr, (n, p)), device="cuda")
    if zero_cols is not None and np.any(zero_cols):
        R[:, torch.as_tensor(np.where(zero_cols)[0], device="cuda")] = 0.0
    b = int(min(n, r + oversample) if shard else min(min(n, p), r + oversample))   # (same block width on every rank)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed + (dist_info()[0] if shard else 0))
    V0 = torch.randn(p, b, dtype=torch.float64, device="cuda", generator=gen)

    def orth_chol(X, rows_sharded=False):
        # CholeskyQR2: two rounds of X <- X R^-1 with R^T R = X^T X (a b x b Gram, Cholesky and a
        # triangular solve -- three small kernels instead of cuSOLVER's Householder panel, which
        # took 0.3 ms per call on these tall-skinny blocks and half of this routine's time).  With the
        # rows of X split over ranks only the Gram matrix needs a reduction.
        for _ in range(2):
            G = X.T @ X
            if rows_sharded:
                G = red(G)
            Lc = torch.linalg.cholesky_ex(G)[0]
            X = torch.linalg.solve_triangular(Lc, X.T, upper=False).T
        return X

    def orth_qr(X, rows_sharded=False):
        if rows_sharded:
            return orth_chol(X, True)
        return torch.linalg.qr(X)[0]

    def iterate(orth):
        V = V0
        Ub = S = Vr = None
        for it in range(max_iter):
            Q = orth(red(R @ V))                   # (n, b), replicated
            V = orth(R.T @ Q, shard)               # (p, b), rows = this rank's genes
            if it % 4 == 3 or it == max_iter - 1:
                Bm = red(R @ V)                     # (n, b): R restricted to span(V)
                Ub, S, Wt = torch.linalg.svd(Bm, full_matrices=False)
                # Ritz residual of the leading r right vectors: || R^T u - s v ||
                Vr = V @ Wt[:r].T                   # (p, r)
                res2 = red(((R.T @ Ub[:, :r] - Vr * S[:r]) ** 2).sum(dim=0))
                res = torch.sqrt(res2).max().item()
                if not np.isfinite(res):
                    return None
                if res <= tol * S[0].item():
                    break
        return Ub, S, Vr

    out = iterate(orth_chol)
    if out is None:        # Gram factorisation broke down (rank-deficient block): Householder QR
        out = iterate(orth_qr)
    Ub, S, Vr = out
    # scipy's svds returns the k triplets in ASCENDING order of the singular values; the latent
    # columns of X_U / B_Gamma inherit that order, so it is kept for identical outputs
    u = Ub[:, :r].cpu().numpy()[:, ::-1].copy()
    s = S[:r].cpu().numpy()[::-1].copy()
    vt = Vr.T.cpu().numpy()[::-1].copy()
    u, vt = fix_signs(u, vt)
    return u, s, vt


def lowrank_product_svd(U, G, sharded=False):
    """SVD of ``E = U @ G.T`` (n x p, rank r) from its factors ``U (n, r)``, ``G (p, r)``.
    ``sharded``: ``G`` holds this rank's rows (genes); its orthogonalisation goes through the
    all-reduced r x r Gram matrix (Cholesky-QR) and ``vt`` comes back with this rank's columns.

    Replaces ``svds(E, k=r)`` of ``causarray/gcate_opt.py:330-331`` without
    forming ``E``: thin QR of both factors and an r x r SVD.
    Returns ``u (n, r)``, ``s (r,)`` ascending, ``vt (r, p)``.
    """
    Qa, Ra = np.linalg.qr(U)
    if sharded:
        from .parallel import allreduce_sum_np
        Rb = np.linalg.cholesky(allreduce_sum_np(G.T @ G)).T      # G = Qb Rb, Rb upper triangular
        Qb = np.linalg.solve(Rb.T, G.T).T
    else:
        Qb, Rb = np.linalg.qr(G)
    uu, s, vvt = np.linalg.svd(Ra @ Rb.T)
    u = (Qa @ uu)[:, ::-1].copy()        # ascending singular values, as svds returns them
    vt = (vvt @ Qb.T)[::-1].copy()
    s = s[::-1].copy()
    u, vt = fix_signs(u, vt)
    return u, s, vt
