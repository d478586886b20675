"""How much of an NB fit is the tail of a few slow genes?  (not a benchmark)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import bench
from causarray_b200 import _lib as L
from causarray_b200.device import GlmDevice

b = bench.make_batch(7)
Y, X, A = b["Y"], b["X"], b["A"]
n, p = Y.shape
rng = np.random.default_rng(0)
U = rng.standard_normal((n, 10)) * 0.05
Xf = np.c_[X, U, A]
off = np.log(np.maximum(Y.sum(1), 1) / np.mean(Y.sum(1)))
dev = GlmDevice(n, p)
dev.load_counts(Y)
for mi in (24, 100, 400):
    dev.fit("nb", Xf, off, b["disp"], maxiter=mi, d_guard=11)
    L.profile_enable(True); L.profile_reset()
    t0 = time.perf_counter()
    B, dv, it, st = dev.fit("nb", Xf, off, b["disp"], maxiter=mi, d_guard=11)
    dt = time.perf_counter() - t0
    prof = L.profile_snapshot(); L.profile_enable(False)
    print("maxiter", mi, "fit %.1f ms" % (dt * 1e3), "max it", int(it.max()),
          {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]})
j = int(np.argmax(it))
print("slowest gene", j, "n_iter", int(it[j]), "status", int(st[j]), "max|beta|", float(np.abs(B[j]).max()), "dev", float(dv[j]),
      "mean count", float(Y[:, j].mean()), "nonzero cells", int((Y[:, j] > 0).sum()))
for mi in (30, 60, 100, 200):
    B2, dv2, it2, st2 = dev.fit("nb", Xf, off, b["disp"], maxiter=mi, d_guard=11)
    print("  maxiter", mi, "dev %.12f" % dv2[j], "beta[:4]", B2[j, :4], "max|beta|", float(np.abs(B2[j]).max()), "status", int(st2[j]))
