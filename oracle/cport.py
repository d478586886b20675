"""ctypes wrapper of ``oracle/gcate_oracle.c`` (TEST INFRASTRUCTURE / timed CPU baseline).

``update_with_mask`` has the signature of the reference kernel
(``causarray/gcate_opt.py:148-152``) so it can be dropped into
``oracle.altmin.alter_min_loop`` in place of the numpy restatement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libgcate_oracle.so")
_lib = None


def build():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(_HERE, "gcate_oracle.c")):
        os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", SO,
                               os.path.join(_HERE, "gcate_oracle.c"), "-lm"])
    return SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(SO)
        D = ctypes.POINTER(ctypes.c_double)
        U8 = ctypes.POINTER(ctypes.c_uint8)
        lib.gcate_update_with_mask.restype = ctypes.c_double
        lib.gcate_update_with_mask.argtypes = (
            [ctypes.c_int] * 5 + [D, D, D, D, D, D, ctypes.c_int, D, ctypes.c_int, ctypes.c_int, D, D, D]
            + [ctypes.c_double] * 4 + [ctypes.c_int, ctypes.c_double, U8, U8, D, ctypes.POINTER(ctypes.c_longlong)])
        lib.gcate_oracle_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def threads():
    return load().gcate_oracle_threads()


def set_threads(n):
    """Number of OpenMP threads of the C restatement (torchrun exports OMP_NUM_THREADS=1)."""
    lib = load()
    lib.gcate_oracle_set_threads.argtypes = [ctypes.c_int]
    lib.gcate_oracle_set_threads.restype = None
    lib.gcate_oracle_set_threads(int(n))
    return lib.gcate_oracle_threads()


def _d(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Problem:
    """Keeps the transposed copies of Y and Ys so that repeated iterations do not redo them."""

    def __init__(self, Y, Ys):
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.Ys = np.ascontiguousarray(Ys, dtype=np.float64)
        self.Yt = np.ascontiguousarray(self.Y.T)
        self.Yst = np.ascontiguousarray(self.Ys.T)
        self.evals = (ctypes.c_longlong * 2)(0, 0)


def update_with_mask(Y, A, B, d, lam, P1, P2, family, nu, Ys, thres_disp, a, C, alpha, beta, max_iters, tol,
                     gene_active, cell_active, alpha_gene, stats=None, _prob=None):
    lib = load()
    prob = _prob if _prob is not None else Problem(Y, Ys)
    n, p = prob.Y.shape
    k = A.shape[1]
    assert A.flags["C_CONTIGUOUS"] and B.flags["C_CONTIGUOUS"] and A.dtype == np.float64 and B.dtype == np.float64
    lam_arr = None
    if lam is not None and np.ndim(lam) == 2 and np.any(lam != 0):
        lam_arr = np.ascontiguousarray(lam, dtype=np.float64)
        assert lam_arr.shape == (p, a)
    P1c = None if P1 is None else np.ascontiguousarray(P1, dtype=np.float64)
    P2c = None if P2 is None else np.ascontiguousarray(P2, dtype=np.float64)
    nu_c = np.ascontiguousarray(np.ravel(nu), dtype=np.float64)
    ga = None if gene_active is None else np.ascontiguousarray(gene_active, dtype=np.uint8)
    ca = None if cell_active is None else np.ascontiguousarray(cell_active, dtype=np.uint8)
    ag = None if alpha_gene is None else np.ascontiguousarray(alpha_gene, dtype=np.float64)
    u8 = lambda x: None if x is None else x.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    f = lib.gcate_update_with_mask(
        n, p, k, int(d), int(a), _d(prob.Y), _d(prob.Yt), _d(A), _d(B), _d(lam_arr), _d(P1c),
        0 if P1c is None else P1c.shape[1], _d(P2c), 0 if P2c is None else P2c.shape[1],
        0 if family == "poisson" else 1, _d(nu_c), _d(prob.Ys), _d(prob.Yst), float(thres_disp), float(C),
        float(alpha), float(beta), int(max_iters), float(tol), u8(ga), u8(ca), _d(ag), prob.evals)
    if stats is not None:
        stats["cell_evals"] = int(prob.evals[0])
        stats["gene_evals"] = int(prob.evals[1])
    return f, A, B
