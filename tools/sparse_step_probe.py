"""End-to-end time of one benchmark batch given dense vs scipy.sparse counts (not a benchmark)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import scipy.sparse as sp
import bench

b = bench.make_batch(7)
Yd = b["Y"]
Ys = sp.csr_matrix(Yd)
print("density %.3f, dense %.0f MB, csr %.0f MB" % (Ys.nnz / Yd.size, Yd.nbytes / 1e6,
                                                     (Ys.data.nbytes + Ys.indices.nbytes + Ys.indptr.nbytes) / 1e6))
for name, Y in (("dense", Yd), ("csr", Ys), ("dense", Yd), ("csr", Ys)):
    bb = dict(b, Y=Y)
    bench.gpu_step(bb)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        df, st = bench.gpu_step(bb)
        ts.append(time.perf_counter() - t0)
    print(name, "step %.1f ms (min of 3)" % (min(ts) * 1e3), "n_sig", st["n_sig"])
t0 = time.perf_counter(); Ys.toarray(); print("host toarray alone: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
