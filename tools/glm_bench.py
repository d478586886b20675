"""Time the batched IRLS kernels alone on a bench-shaped batch (not a benchmark)."""
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
for fam, nu in (("poisson", None), ("nb", b["disp"])):
    dev.fit(fam, Xf, off, nu, maxiter=100, d_guard=11)
    L.profile_enable(True); L.profile_reset()
    t0 = time.perf_counter()
    B, dv, it, st = dev.fit(fam, Xf, off, nu, maxiter=100, d_guard=11)
    dt = time.perf_counter() - t0
    prof = L.profile_snapshot()
    L.profile_enable(False)
    print(fam, "fit %.1f ms" % (dt * 1e3), "gene-iters", int(it.sum()), "max it", int(it.max()), "bad", int((st != 0).sum()),
          {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]})
    h = np.bincount(np.minimum(it, 100))
    print("   iterations histogram:", {int(k): int(v) for k, v in enumerate(h) if v})
