"""Timing of the initial-SVD phase pieces on a benchmark batch (not a benchmark)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import torch
import bench
from causarray_b200 import glm as cglm, svd as csvd
from causarray_b200.optim import BatchWorkspace
from causarray_b200.prep import comp_size_factor

b = bench.make_batch(7)
Y, X, A = b["Y"], b["X"], b["A"]
n, p = Y.shape
Xf = np.c_[X, A]
ws = BatchWorkspace.upload(Y, "nb", Xf.shape[1], 10)
sf = ws.gcate.size_factors()
ws.prepare(sf, b["disp"])
Bx, fit, _, _, _ = cglm.fit_glm(Y, Xf, offset=np.log(sf), family="nb", disp_glm=b["disp"], maxiter=100, _glm=ws.glm, _lazy=True)
os.environ["CA_SVD_DEBUG"] = "1"
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ptrs = ws.glm.resid_device_ptrs()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    u, s, vt = csvd.topr_svd_native(ws.glm, n, p, 10)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ptr = ws.glm.resid_device_ptr()
    u2, s2, vt2 = csvd.topr_svd_device(ptr, n, p, 10)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("resid2 %.2f ms, native (incl. resid2) %.2f ms, torch (incl. resid) %.2f ms; s rel diff %.2e, |u|diff %.2e"
          % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, np.abs(s / s2 - 1).max(), np.abs(u - u2).max()))
