"""Run under torchrun with W ranks: fit_gcate with gene_shard=True (GLM initialisation, sharded SVD,
both stages of the alternating minimisation with the genes split over the ranks) against the same
fit on one GPU.  Exit code 0 on success."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import causarray_b200 as cb
    from causarray_b200 import _lib as L
    from causarray_b200.synth import simulate_screen
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    L.load()
    L.check(L._lib.ca_set_device(local))
    dist.init_process_group("gloo")
    warnings.simplefilter("ignore")
    sim = simulate_screen(n_ctrl=220, n_perts=3, cells_per_pert=60, p=301, r=2, d_cov=1, family="nb", seed=17,
                          base_lo=0.0, base_hi=2.0)
    Y, X, A = sim["Y"], sim["X"], sim["A"]
    kw = dict(family="nb", disp_glm=sim["disp"], kwargs_es_1=dict(max_iters=25), kwargs_es_2=dict(max_iters=25))
    r1, r2 = cb.fit_gcate(Y.copy(), X, A, 2, **kw)                      # whole fit on this rank's GPU
    ws = r2.pop("_ws")
    W = np.c_[X, r2["U"]]
    off = np.log(r2["kwargs_glm"]["size_factor"])
    df_ref, _ = cb.LFC(Y.copy(), W, A, W_A=W, family="nb", offset=off, disp_glm=sim["disp"], _glm=ws.glm, _trusted=True)
    ws.close()
    s1, s2 = cb.fit_gcate(Y.copy(), X, A, 2, gene_shard=True, **kw)     # genes split over the ranks
    ws = s2.pop("_ws")
    Ws = np.c_[X, s2["U"]]
    df_sh, _ = cb.LFC(Y.copy(), Ws, A, W_A=Ws, family="nb", offset=np.log(s2["kwargs_glm"]["size_factor"]),
                      disp_glm=sim["disp"], _glm=ws.glm, _trusted=True, gene_shard=True)
    ws.close()
    lo, hi = s2["gene_slice"]
    ok = True

    def check(name, got, want, tol):
        nonlocal ok
        err = float(np.max(np.abs(np.asarray(got) - np.asarray(want))))
        good = err <= tol
        ok &= good
        print("rank %d genes [%d, %d) %-10s max|diff| %.3e %s" % (rank, lo, hi, name, err, "ok" if good else "FAIL"), flush=True)

    ok &= len(s1["hist"]) == len(r1["hist"]) and len(s2["hist"]) == len(r2["hist"])
    print("rank", rank, "history lengths", len(s1["hist"]), len(r1["hist"]), len(s2["hist"]), len(r2["hist"]), flush=True)
    if ok:
        check("s1 hist", np.array(s1["hist"]) / np.array(r1["hist"]) - 1, 0.0, 1e-7)
        check("s2 hist", np.array(s2["hist"]) / np.array(r2["hist"]) - 1, 0.0, 1e-7)
    check("s1 X_U", s1["X_U"], r1["X_U"], 2e-5)
    check("s1 B", s1["B_Gamma"], r1["B_Gamma"][lo:hi], 2e-5)
    check("s2 X_U", s2["X_U"], r2["X_U"], 2e-5)
    check("s2 B", s2["B_Gamma"], r2["B_Gamma"][lo:hi], 2e-5)
    ok &= len(df_sh) == len(df_ref) and list(df_sh.columns) == list(df_ref.columns)
    for c in ("tau", "std", "pvalue", "padj"):
        a_, b_ = df_sh[c].to_numpy().astype(float), df_ref[c].to_numpy().astype(float)
        fin = np.isfinite(b_)
        ok &= bool(np.array_equal(np.isfinite(a_), fin))
        check("LFC " + c, a_[fin] / np.where(np.abs(b_[fin]) > 1e-3, b_[fin], 1.0), b_[fin] / np.where(np.abs(b_[fin]) > 1e-3, b_[fin], 1.0), 1e-5)
    ok &= bool(np.array_equal(df_sh["padj"].to_numpy() < 0.05, df_ref["padj"].to_numpy() < 0.05))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
