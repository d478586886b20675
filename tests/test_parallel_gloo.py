"""CPU, world_size 2 over gloo: the batch-sharding logic of gcate_lfc_batch."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch.multiprocessing as mp

from causarray_b200 import parallel as par


def test_owned_batches_partition():
    for nb in (1, 2, 7, 14):
        for w in (1, 2, 4, 8):
            owned = [par.owned_batches(nb, r, w) for r in range(w)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(nb))
            sizes = [len(o) for o in owned]
            assert max(sizes) - min(sizes) <= 1
            for r in range(w):
                assert par.foreign_batches(nb, r, w) == set(range(nb)) - set(owned[r])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert par.dist_info() == (rank, world)
    mine = {i: pd.DataFrame({"tau": np.full(3, float(i)), "batch": i}) for i in par.owned_batches(5, rank, world)}
    merged = par.gather_frames(mine)
    if rank == 0:
        q.put(sorted(merged))
        q.put(float(pd.concat([merged[i] for i in sorted(merged)])["tau"].sum()))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_frames_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys = q.get(timeout=120)
    total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert keys == [0, 1, 2, 3, 4]
    assert total == 3 * (0 + 1 + 2 + 3 + 4)


def test_gene_bounds_cover():
    for p in (1, 7, 301, 8000):
        for w in (1, 2, 3, 8):
            b = par.gene_bounds(p, w)
            assert b[0][0] == 0 and b[-1][1] == p and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker_sharded_linalg(rank, world, port, q):
    """Host pieces of the gene-sharded initialisation: the low-rank product SVD and the Cholesky-QR of
    Gamma-hat with the genes (rows) split over two ranks against the unsharded results."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from causarray_b200.svd import lowrank_product_svd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    n, p, r = 40, 61, 3
    U, G = rng.standard_normal((n, r)), rng.standard_normal((p, r))
    lo, hi = par.gene_bounds(p, world)[rank]
    u_ref, s_ref, vt_ref = lowrank_product_svd(U, G)
    u, s, vt = lowrank_product_svd(U, G[lo:hi], sharded=True)
    ok = np.allclose(s, s_ref, rtol=1e-12) and np.allclose(u, u_ref, atol=1e-10) and np.allclose(vt, vt_ref[:, lo:hi], atol=1e-10)
    tot = par.allreduce_sum_np(np.array([float(hi - lo)]))
    ok = ok and tot[0] == p
    t = par.allreduce_sum(torch.ones(3, dtype=torch.float64) * (rank + 1))
    ok = ok and bool((t == 3.0).all())
    blob = par.broadcast_bytes(b"id-from-rank-0" if rank == 0 else None)
    ok = ok and blob == b"id-from-rank-0"
    # per-gene result blocks of the sharded LFC, gathered along the gene axis in rank order
    full = rng.standard_normal((4, p))
    got = par.allgather_concat(full[:, lo:hi], axis=1)
    ok = ok and np.array_equal(got, full)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_linalg_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_sharded_linalg, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def _worker_sharded_lfc(rank, world, port, q):
    """Host logic of LFC(gene_shard=True) at world size 2: column slicing, the gene-axis gather and the BH
    step over ALL genes -- with the device pieces (outcome GLM, AIPW reduction) replaced by numpy
    stand-ins that depend on the gene's own counts only, as the real ones do."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import warnings
    import torch.distributed as dist
    from causarray_b200 import dr
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    n, p, a = 90, 37, 3
    A = np.zeros((n, a))
    for j in range(a):
        A[30 + 20 * j:50 + 20 * j, j] = 1.0
    Y = rng.poisson(3.0 + 2.0 * rng.random(p)[None, :] + 1.5 * A[:, :1] * (np.arange(p) % 5 == 0)[None, :],
                    size=(n, p)).astype(float)
    W = np.c_[np.ones(n), rng.standard_normal(n)]

    class FakeGlm:
        def __init__(self, n_, p_):
            self.n, self.p = n_, p_

        def load_counts(self, Y_):
            assert Y_.shape == (self.n, self.p)

    def fake_cross_fitting(Y_, A_, X_, X_A_, **kw):
        p_loc = Y_.shape[1]
        pi = np.full(A_.shape, 0.5)
        return dr.YhatFactors(X_, np.zeros((p_loc, X_.shape[1])), np.zeros((p_loc, A_.shape[1])), None), pi, pi

    def fake_aipw_lfc(Y_, W_, Bcov, Btrt, offset, sf, A_, pi, cell_mask=None, glm=None, **kw):
        ctrl = A_.sum(axis=1) == 0
        out = {k: np.zeros((A_.shape[1], Y_.shape[1])) for k in ("tau", "var", "df", "mean0", "mean1")}
        out["estimable"] = np.ones((A_.shape[1], Y_.shape[1]), dtype=bool)
        for j in range(A_.shape[1]):
            y0, y1 = Y_[ctrl], Y_[A_[:, j] == 1]
            out["mean0"][j], out["mean1"][j] = y0.mean(0), y1.mean(0)
            out["tau"][j] = np.log((y1.mean(0) + 1) / (y0.mean(0) + 1))
            out["var"][j] = y0.var(0, ddof=1) / len(y0) / (y0.mean(0) + 1) ** 2 + y1.var(0, ddof=1) / len(y1) / (y1.mean(0) + 1) ** 2
            out["df"][j] = len(y0) + len(y1) - 2.0
        return out

    dr.GlmDevice, dr.cross_fitting, dr.aipw_lfc = FakeGlm, fake_cross_fitting, fake_aipw_lfc
    from causarray_b200 import inference as _inf
    dr.bh_correction_device = _inf.bh_correction_multi     # host form of the same tail (no GPU in this test)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_full, _ = dr.LFC(Y, W, A, family="poisson", offset=np.zeros(n), _trusted=True)
        df_sh, _ = dr.LFC(Y, W, A, family="poisson", offset=np.zeros(n), _trusted=True, gene_shard=True)
    ok = len(df_sh) == a * p and list(df_sh.columns) == list(df_full.columns)
    for c in ("tau", "std", "pvalue", "padj", "mean_control", "mean_treated"):
        ok = ok and np.array_equal(df_sh[c].to_numpy(), df_full[c].to_numpy(), equal_nan=True)
    ok = ok and list(df_sh["gene_names"]) == list(df_full["gene_names"])
    ok = ok and bool((df_full["padj"] < 0.05).sum() > 0)        # the BH step had something to decide
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lfc_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_sharded_lfc, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
