"""Time the alternating-iteration kernels alone on a bench-shaped batch (not a benchmark)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import scipy.linalg
import bench
from causarray_b200 import _lib as L
from causarray_b200.device import GcateDevice

b = bench.make_batch(7)
Y, X, A = b["Y"], b["X"], b["A"]
Xf = np.c_[X, A]
n, p = Y.shape
d, r = Xf.shape[1], 10
rng = np.random.default_rng(0)
dev = GcateDevice(n, p, d, r, "nb", 100.0)
dev.upload_counts(Y, check=False)
sf = dev.size_factors()
dev.prepare(sf, b["disp"])
A0 = np.c_[Xf, rng.standard_normal((n, r)) * 0.05]
B0 = np.c_[np.log(Y.mean(0) + 0.05)[:, None], rng.standard_normal((p, d - 1 + r)) * 0.05]
Q1 = scipy.linalg.qr(Xf, mode="economic")[0]
dev.set_factors(A0, B0)
dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
for it in range(2):
    dev.update(1e3, 0.1, 0.5, 20, 1e-4)
L.profile_enable(True); L.profile_reset()
t0 = time.perf_counter()
for it in range(10):
    f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
dt = time.perf_counter() - t0
prof = L.profile_snapshot()
print(os.environ.get("CAUSARRAY_B200_LIB", "default"), "10 iters %.1f ms" % (dt * 1e3), "f", f,
      {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]})
