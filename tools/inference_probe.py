"""Host-side timing of the inference tail on this box (not a benchmark)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.special as sc
from causarray_b200 import inference as inf

rng = np.random.default_rng(0)
a, p = 15, 8000
t = rng.standard_normal((a, p)) * 2
df = rng.uniform(50, 400, (a, p))


def T(f, n=7):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return "min %.2f med %.2f max %.2f ms" % (min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3, max(ts) * 1e3)


print("cpus", os.cpu_count(), "threads", inf._N_THREADS)
print("bh_correction_multi", T(lambda: inf.bh_correction_multi(t, df=df)))
print("tail_parallel", T(lambda: inf._tail_parallel(t, df)))
print("stdtr 1 thread", T(lambda: sc.stdtr(df, -np.abs(t))))
pv = inf._tail_parallel(t, df)
print("bh rows", T(lambda: inf._bh_rows_block(pv)))
print("nanmedian x2", T(lambda: (np.nanmedian(t, axis=1), np.nanmedian(np.abs(t), axis=1))))
