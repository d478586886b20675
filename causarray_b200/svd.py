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


def topr_svd_device(ptr, n, p, r, zero_cols=None, oversample=20, max_iter=400, tol=1e-11, seed=0):
    """Top-``r`` singular triplets of the (n, p) float64 matrix at device pointer ``ptr``.

    Returns host arrays ``u (n, r)``, ``s (r,)`` (ascending, as ``svds``), ``vt (r, p)``.
    """
    import torch
    if not torch.cuda.is_available():
        from ._lib import CausarrayCudaError
        raise CausarrayCudaError("no CUDA device for the initial SVD")
    R = torch.as_tensor(_DevArray(ptr, (n, p)), device="cuda")
    if zero_cols is not None and np.any(zero_cols):
        R[:, torch.as_tensor(np.where(zero_cols)[0], device="cuda")] = 0.0
    b = int(min(min(n, p), r + oversample))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    V0 = torch.randn(p, b, dtype=torch.float64, device="cuda", generator=gen)

    def orth_chol(X):
        # CholeskyQR2: two rounds of X <- X R^-1 with R^T R = X^T X (a b x b Gram, Cholesky and a
        # triangular solve -- three small kernels instead of cuSOLVER's Householder panel, which
        # took 0.3 ms per call on these tall-skinny blocks and half of this routine's time)
        for _ in range(2):
            Lc = torch.linalg.cholesky_ex(X.T @ X)[0]
            X = torch.linalg.solve_triangular(Lc, X.T, upper=False).T
        return X

    def orth_qr(X):
        return torch.linalg.qr(X)[0]

    def iterate(orth):
        V = V0
        Ub = S = Vr = None
        for it in range(max_iter):
            Q = orth(R @ V)                        # (n, b)
            V = orth(R.T @ Q)                      # (p, b)
            if it % 4 == 3 or it == max_iter - 1:
                Bm = R @ V                          # (n, b): R restricted to span(V)
                Ub, S, Wt = torch.linalg.svd(Bm, full_matrices=False)
                # Ritz residual of the leading r right vectors: || R^T u - s v ||
                Vr = V @ Wt[:r].T                   # (p, r)
                res = torch.linalg.norm(R.T @ Ub[:, :r] - Vr * S[:r], dim=0).max().item()
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


def lowrank_product_svd(U, G):
    """SVD of ``E = U @ G.T`` (n x p, rank r) from its factors ``U (n, r)``, ``G (p, r)``.

    Replaces ``svds(E, k=r)`` of ``causarray/gcate_opt.py:330-331`` without
    forming ``E``: thin QR of both factors and an r x r SVD.
    Returns ``u (n, r)``, ``s (r,)`` ascending, ``vt (r, p)``.
    """
    Qa, Ra = np.linalg.qr(U)
    Qb, Rb = np.linalg.qr(G)
    uu, s, vvt = np.linalg.svd(Ra @ Rb.T)
    u = (Qa @ uu)[:, ::-1].copy()        # ascending singular values, as svds returns them
    vt = (vvt @ Qb.T)[::-1].copy()
    s = s[::-1].copy()
    u, vt = fix_signs(u, vt)
    return u, s, vt
