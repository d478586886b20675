"""Run under torchrun with W ranks: the two-stage alternating minimisation of the alter_min golden
(real-reference outputs) with the GENES split over the ranks; every rank checks its shard of B and the
replicated A / history against the golden.  Exit code 0 on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import scipy.linalg
    import torch
    import torch.distributed as dist
    from causarray_b200 import _lib as L
    from causarray_b200.device import GcateDevice, comm_unique_id
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    L.load()
    L.check(L._lib.ca_set_device(local))
    dist.init_process_group("gloo")
    fam = sys.argv[1] if len(sys.argv) > 1 else "nb"
    g = np.load(os.path.join(ROOT, "tests", "golden", "alter_min_%s.npz" % fam))
    Y, X = g["Y"], g["X"]
    n, p = Y.shape
    d, r, a = X.shape[1], int(g["r"]), int(g["a"])
    edges = np.linspace(0, p, world + 1).astype(int)
    lo, hi = int(edges[rank]), int(edges[rank + 1])
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dev = GcateDevice(n, hi - lo, d, r, fam, 100.0)
    dev.set_shard(rank, world, p, box[0])
    dev.load_counts(np.ascontiguousarray(Y[:, lo:hi]), g["sf"], np.ascontiguousarray(g["disp"][lo:hi]))
    ls = {"alpha": 0.1, "beta": 0.5, "max_iters": 20, "tol": 1e-4, "C": 1e3 if fam == "nb" else 1e5,
          "tol_gene": 1e-4, "tol_cell": 1e-4, "recheck_interval": 10, "sparsity_threshold": 0.5,
          "sparsity_boost": 2.0, "warmup_iters": 0}
    es = {"max_iters": 30, "warmup": 0, "patience": 5, "tolerance": 0.0, "rel_tol": 2e-4}
    Q1 = scipy.linalg.qr(X, mode="economic")[0]
    dev.set_factors(g["A0"], np.ascontiguousarray(g["B0"][lo:hi]))
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    res, hist = dev.alter_min(ls, es, 0.0, d)
    A, B = dev.get_factors()
    ok = True
    if rank == 0:
        print("world %d, cell-step reductions over: %s" % (world, dev.transport()), flush=True)

    def check(name, got, want, tol):
        nonlocal ok
        err = float(np.max(np.abs(np.asarray(got) - np.asarray(want)))) if np.size(want) else 0.0
        good = err <= tol
        ok &= good
        print("rank %d %-10s max|diff| %.3e %s" % (rank, name, err, "ok" if good else "FAIL"), flush=True)

    ok &= res.n_iter == int(g["s1_n_iter"])
    print("rank", rank, "stage 1 n_iter", res.n_iter, int(g["s1_n_iter"]), flush=True)
    check("s1 hist", hist / g["s1_hist"][:len(hist)] - 1.0, 0.0, 1e-9)
    check("s1 A", A, g["s1_A"], 1e-7)
    check("s1 B", B, g["s1_B"][lo:hi], 1e-7)
    dev.set_factors(g["s1_A"], np.ascontiguousarray(g["s1_B"][lo:hi]))
    dev.set_stage(P1=None, P2=np.ascontiguousarray(g["Q2"][lo:hi]), lam=None, a=a)
    res2, hist2 = dev.alter_min(ls, es, 0.05, a)
    A2, B2 = dev.get_factors()
    ok &= res2.n_iter == int(g["s2_n_iter"])
    print("rank", rank, "stage 2 n_iter", res2.n_iter, int(g["s2_n_iter"]), flush=True)
    check("s2 hist", hist2 / g["s2_hist"][:len(hist2)] - 1.0, 0.0, 1e-9)
    check("s2 A", A2, g["s2_A"], 1e-7)
    check("s2 B", B2, g["s2_B"][lo:hi], 1e-7)
    dev.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
