"""Which GLM fits does a benchmark step run, and how long is each?  (not a benchmark)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import bench
from causarray_b200 import glm as cglm
from causarray_b200 import _lib as L

orig = cglm.GlmDevice.fit if hasattr(cglm, "GlmDevice") else None
from causarray_b200.device import GlmDevice
_fit = GlmDevice.fit
SAVED = []


def fit(self, family, X, *a, **k):
    L.profile_enable(True); L.profile_reset()
    t0 = time.perf_counter()
    out = _fit(self, family, X, *a, **k)
    dt = time.perf_counter() - t0
    prof = L.profile_snapshot(); L.profile_enable(False)
    it, st = out[2], out[3]
    SAVED.append(("fit_%s_kx%d" % (family, X.shape[1]), np.asarray(it).copy(), np.asarray(st).copy(), np.asarray(out[0]).copy()))
    print("  fit %-7s kx=%d maxiter=%s ridge=%s: %.1f ms wall; kernels %s; max it %d, gene-iters %d, status!=0: %d"
          % (family, X.shape[1], k.get("maxiter"), k.get("ridge") is not None, dt * 1e3,
             {kk: (round(v[0], 2), v[1]) for kk, v in prof.items() if v[1]}, int(it.max()), int(it.sum()), int((st != 0).sum())))
    return out


GlmDevice.fit = fit
_refit = GlmDevice.refit_l1


def refit(self, family, X, l1, gene_mask, **k):
    L.profile_enable(True); L.profile_reset()
    t0 = time.perf_counter()
    out = _refit(self, family, X, l1, gene_mask, **k)
    dt = time.perf_counter() - t0
    prof = L.profile_snapshot(); L.profile_enable(False)
    it, st = out[2], out[3]
    m = np.asarray(gene_mask, bool)
    SAVED.append(("refit_kx%d" % X.shape[1], np.asarray(it).copy(), m.copy(), np.asarray(out[0]).copy()))
    print("  lasso refit of %d genes: %.1f ms wall; kernels %s; iterations histogram %s, status!=0: %d"
          % (int(m.sum()), dt * 1e3, {kk: (round(v[0], 2), v[1]) for kk, v in prof.items() if v[1]},
             {int(a): int(b) for a, b in enumerate(np.bincount(it[m])) if b}, int((st[m] != 0).sum())))
    return out


GlmDevice.refit_l1 = refit
b = bench.make_batch(7)
bench.gpu_step(b)
print("---- second step")
bench.gpu_step(b)

os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/glm_iters.npz", **{"%02d_%s_%s" % (i, nm, k): v for i, (nm, a, b, c) in enumerate(SAVED) for k, v in (("it", a), ("st", b), ("B", c))})
