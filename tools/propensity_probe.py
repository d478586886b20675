"""Host-side timing of the propensity fits on this box (not a benchmark)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import bench
from causarray_b200 import dr

b = bench.make_batch(7)
A, X_A = b["A"], b["X_A"]
rng = np.random.default_rng(0)
W_A = np.c_[X_A, rng.standard_normal((A.shape[0], 10)) * 0.1]
for rep in range(3):
    t0 = time.perf_counter()
    pi = dr.estimate_propensity_scores(A, W_A)
    print("threads: %.1f ms" % ((time.perf_counter() - t0) * 1e3))

from joblib import Parallel, delayed
from sklearn.linear_model import LogisticRegression


def one(x, y, xt):
    return LogisticRegression(fit_intercept=False, C=1.0, class_weight="balanced", random_state=0).fit(x, y).predict_proba(xt)[:, 1]


ctrl = A.sum(1) == 0
par = Parallel(n_jobs=8, backend="loky")
for rep in range(4):
    t0 = time.perf_counter()
    tasks = []
    for j in range(A.shape[1]):
        el = ctrl | (A[:, j] == 1)
        tasks.append(delayed(one)(W_A[el], A[el, j], W_A))
    out = par(tasks)
    print("loky 8 procs: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
print("max diff", max(np.abs(out[j] - pi[:, j]).max() for j in range(A.shape[1])))
t0 = time.perf_counter()
for j in range(A.shape[1]):
    el = ctrl | (A[:, j] == 1)
    one(W_A[el], A[el, j], W_A)
print("sequential: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
