"""Drive the REAL reference (numba ``alter_min``, its own ``fit_gcate`` / ``LFC`` orchestration)
on one perturbation batch -- TEST INFRASTRUCTURE: the CPU side of the full-size parity tools and
the ``--impl reference`` arm of ``bench.py``.

The reference is imported through ``oracle/ref_shim.py`` (``/root/reference`` in the build
container, the staged copy ``oracle/_ref`` on the GPU box).  Its per-gene GLM calls
(``sm.GLM(...).fit``, ``causarray/gcate_glm.py:204``) are served by the statsmodels stand-in of
``oracle/shims`` -- the one piece of the reference's arithmetic that lives in an absent third-party
package (parity unpinned for it, see ``oracle/glm.py``).  Everything else -- size factors, the GLM
+ SVD initialisation, the numba kernels of the alternating minimisation, propensities, AIPW, the
estimand, BH -- is the reference's own code, unmodified.

``scipy.sparse.linalg.svds`` is unseeded in the reference (``gcate_opt.py:312,331``);
``deterministic_svds`` pins the ARPACK start vector and the signs so that two runs (and the device
path) see the same initialisation.
"""
import atexit
import contextlib
import os
import shutil
import tempfile
import time

import numpy as np

from . import ref_shim


def _svds_det_factory(orig):
    def svds_det(M, k, **kw):
        v0 = np.ones(min(M.shape)) / np.sqrt(min(M.shape))
        u, s, vt = orig(M, k=k, v0=v0, tol=1e-14)
        for i in range(k):
            j = np.argmax(np.abs(u[:, i]))
            if u[j, i] < 0:
                u[:, i] *= -1
                vt[i] *= -1
        return u, s, vt
    return svds_det


@contextlib.contextmanager
def deterministic_svds():
    ref_shim.import_reference()
    import scipy.sparse.linalg as sla
    import causarray.gcate_opt as go
    import causarray.gcate as gc
    orig = sla.svds
    det = _svds_det_factory(orig)
    saved = (go.svds, go.sp.sparse.linalg.svds, getattr(gc, "svds", None))
    go.svds = det
    go.sp.sparse.linalg.svds = det
    if saved[2] is not None:
        gc.svds = det
    try:
        yield
    finally:
        go.svds = saved[0]
        go.sp.sparse.linalg.svds = saved[1]
        if saved[2] is not None:
            gc.svds = saved[2]


# The counter directory is fixed when this module is imported -- before joblib spawns its (reused)
# worker processes, which inherit the environment once, at spawn time.
_ITER_DIR = tempfile.mkdtemp(prefix="ca_iters_")
os.environ["CA_ORACLE_ITER_DIR"] = _ITER_DIR
atexit.register(shutil.rmtree, _ITER_DIR, True)


def _drain_iteration_files():
    tot = fits = 0
    for f in os.listdir(_ITER_DIR):
        path = os.path.join(_ITER_DIR, f)
        with open(path) as fh:
            for line in fh:
                tot += int(line)
                fits += 1
        open(path, "w").close()
    return tot, fits


@contextlib.contextmanager
def count_irls_iterations():
    """Sum of the IRLS iteration counts of every ``sm.GLM.fit`` call made inside the block, over
    all worker processes (``oracle/glm.py:_count_iterations``).  Yields a dict filled on exit."""
    out = {}
    _drain_iteration_files()
    try:
        yield out
    finally:
        out["irls_gene_iters"], out["glm_fits"] = _drain_iteration_files()


def warm_up_jit(family="nb"):
    """Compile the numba kernels on a 40 x 60 problem (37-60 s) so that timed runs exclude the JIT
    (SURVEY 8d).  Both stage signatures (P1 array / P2 array) are exercised."""
    ca = ref_shim.import_reference()
    import causarray.gcate_glm as gg
    from causarray_b200.synth import simulate_screen
    sim = simulate_screen(n_ctrl=30, n_perts=2, cells_per_pert=5, p=60, r=2, d_cov=1, family=family, seed=3,
                          base_lo=0.0, base_hi=2.0)
    t0 = time.perf_counter()
    with deterministic_svds(), gg._backend_override("original"):
        ca.fit_gcate(sim["Y"].copy(), sim["X"], sim["A"], 2, family=family, disp_glm=sim["disp"].copy(),
                     kwargs_es_1=dict(max_iters=3), kwargs_es_2=dict(max_iters=3), n_jobs=1)
    return time.perf_counter() - t0


def reference_fit_gcate_lfc(Y, X, A, X_A, r, family="nb", disp=None, es1=None, es2=None, usevar="unequal",
                            n_jobs=None):
    """``fit_gcate`` + ``LFC`` of one batch exactly as ``gcate_lfc_batch`` calls them
    (``causarray/DR_learner.py:800-840``), through the reference's own code.  Returns
    ``(res1, res2, df, estimation, stats)``; ``stats['cell_gene_updates']`` follows bench.py's metric."""
    import pandas as pd
    ca = ref_shim.import_reference()
    import causarray.gcate_glm as gg
    stats = {}
    kw_g = {} if n_jobs is None else dict(n_jobs=n_jobs)
    Yf = np.asarray(Y, dtype=np.float64)
    t0 = time.perf_counter()
    with deterministic_svds(), gg._backend_override("original"), count_irls_iterations() as cnt:
        res1, res2 = ca.fit_gcate(Yf.copy(), X, A, r, family=family, disp_glm=None if disp is None else disp.copy(),
                                  offset=True, kwargs_es_1=dict(es1 or {}), kwargs_es_2=dict(es2 or {}), **kw_g)
        stats["t_gcate"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        U = res2["U"]
        off = np.log(res2["kwargs_glm"]["size_factor"])
        df, est = ca.LFC(Yf.copy(), np.c_[X, U], pd.DataFrame(A), np.c_[X_A, U], family=family, offset=off,
                         usevar=usevar, **kw_g)
        stats["t_lfc"] = time.perf_counter() - t1
    stats["t_total"] = time.perf_counter() - t0
    n, p = Yf.shape
    stats["alt_iters"] = (len(res1["hist"]) - 1) + (len(res2["hist"]) - 1)
    stats["irls_gene_iters"] = cnt["irls_gene_iters"]
    stats["glm_fits"] = cnt["glm_fits"]
    stats["cell_gene_updates"] = n * p * stats["alt_iters"] + n * stats["irls_gene_iters"]
    return res1, res2, df, est, stats
