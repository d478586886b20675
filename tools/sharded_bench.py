"""Strong scaling of the alternating iteration with the GENES split over the ranks (torchrun; not the
contract benchmark).  Every rank builds the same benchmark-shaped batch, keeps its gene shard and runs
the same stage-1 iterations; prints ms per alternating iteration (max over ranks)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import scipy.linalg
import torch
import torch.distributed as dist
import bench
from causarray_b200 import _lib as L
from causarray_b200.device import GcateDevice, comm_unique_id

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
L.load(); L.check(L._lib.ca_set_device(local))
if world > 1:
    dist.init_process_group("gloo")
if len(sys.argv) > 1:      # optional scale factor on the number of cells
    k = int(sys.argv[1])
    bench.WORKLOAD["n_ctrl"] *= k
    bench.WORKLOAD["cells_per_pert"] *= k
b = bench.make_batch(7)
Y, X, A = b["Y"], b["X"], b["A"]
Xf = np.c_[X, A]
n, p = Y.shape
d, r = Xf.shape[1], 10
rng = np.random.default_rng(0)
A0 = np.c_[Xf, rng.standard_normal((n, r)) * 0.05]
B0 = np.c_[np.log(Y.mean(0) + 0.05)[:, None], rng.standard_normal((p, d - 1 + r)) * 0.05]
Q1 = scipy.linalg.qr(Xf, mode="economic")[0]
sf = np.exp(np.log(np.maximum(Y.sum(1), 1)) - np.mean(np.log(np.maximum(Y.sum(1), 1))))
edges = np.linspace(0, p, world + 1).astype(int)
lo, hi = int(edges[rank]), int(edges[rank + 1])
dev = GcateDevice(n, hi - lo, d, r, "nb", 100.0)
if world > 1:
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dev.set_shard(rank, world, p, box[0])
dev.load_counts(np.ascontiguousarray(Y[:, lo:hi]), sf, np.ascontiguousarray(b["disp"][lo:hi]))
dev.set_factors(A0, np.ascontiguousarray(B0[lo:hi]))
dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
for it in range(3):
    f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
iters = 10
for it in range(iters):
    f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters * 1e3
if world > 1:
    t = torch.tensor([dt]); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
if rank == 0:
    print("genes sharded over %d GPU(s): %.3f ms per alternating iteration (n=%d, p=%d, k=%d), f=%.9f" % (world, dt, n, p, d + r, f))
if world > 1:
    dist.destroy_process_group()
