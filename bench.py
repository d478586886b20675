#!/usr/bin/env python
"""bench.py -- GCATE + doubly-robust LFC throughput on the Replogle-K562-shaped workload.

A "step" is one perturbation batch of ``gcate_lfc_batch`` (BASELINE.json config 3): 2,000 control
cells + 15 perturbations x 133 cells (the ``max_cells=2000`` cap), p = 8,000 genes, r = 10 latent
factors, NB family; i.e. ``fit_gcate`` (size factors, 2 GLM initialisations + SVD, two-stage
alternating minimisation) followed by ``LFC`` (propensities, Poisson + NB outcome GLMs, AIPW
reductions, BH) through the public Python API.

Metric: cell*gene updates per second = n*p*(alternating iterations, both stages) + n*(sum over
genes and GLM fits of IRLS iterations), divided by the step time.

  value  -- counts already resident in HBM when the timed region starts
  e2e    -- host numpy arrays in, result DataFrame out (H2D / D2H inside the timed region)

``--impl reference`` times the CPU restatement of the same step (oracle/pipeline.py: C/OpenMP
alternating iteration + numpy IRLS over all host cores) on a bounded gene subsample.
"""
import argparse
import json
import os

# Keep stdout to the single JSON line: NCCL prints its version banner there and libraries may log.
# Everything written to fd 1 while the benchmark runs goes to stderr; the JSON line is written to the
# saved descriptor at the end (emit_json).
_REAL_STDOUT = None


def capture_stdout():
    """Called by main(): from here on fd 1 is stderr; emit_json writes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)

import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="replogle_k562_gcate_lfc_batch", n_ctrl=2000, n_perts=15, cells_per_pert=133, p=8000, r=10,
                family="nb", es_max_iters=30, es_rel_tol=2e-4, usevar="unequal",
                counts_dtype="int32")     # host storage type of the UMI counts


def make_batch(seed, p=None, cells_frac=1.0):
    """One synthetic batch of the workload shape; ``p`` / ``cells_frac`` shrink it for the bounded
    samples of the CPU arms (fewer genes; control and per-perturbation cell counts scaled together)."""
    from causarray_b200.synth import simulate_screen
    from causarray_b200.prep import prep_causarray_data
    w = WORKLOAD
    n_ctrl = max(50, int(round(w["n_ctrl"] * cells_frac)))
    cpp = max(10, int(round(w["cells_per_pert"] * cells_frac)))
    sim = simulate_screen(n_ctrl=n_ctrl, n_perts=w["n_perts"], cells_per_pert=cpp,
                          p=p or w["p"], r=w["r"], d_cov=1, family=w["family"], seed=seed, base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    _, _, X, X_A = prep_causarray_data(Y, A)
    return dict(Y=np.ascontiguousarray(Y.astype(w["counts_dtype"])), X=X, A=A, X_A=X_A, disp=sim["disp"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        # under load = upper half of the samples (the sampler also sees host-side phases)
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def gpu_step(batch, resident_ws=None):
    """One batch through the public API.  Returns (DataFrame, stats)."""
    import pandas as pd
    import causarray_b200 as cb
    from causarray_b200 import glm as cglm
    w = WORKLOAD
    es = dict(rel_tol=w["es_rel_tol"], max_iters=w["es_max_iters"])
    cglm.COUNTERS["irls_gene_iters"] = 0
    cglm.COUNTERS["gram_flops"] = 0.0
    cglm.COUNTERS["gram_flops_structured"] = 0.0
    Y, X, A, X_A = batch["Y"], batch["X"], batch["A"], batch["X_A"]
    res1, res2 = cb.fit_gcate(Y, X, A, w["r"], family=w["family"], disp_glm=batch["disp"], offset=True,
                              kwargs_es_1=dict(es), kwargs_es_2=dict(es), _ws=resident_ws)
    ws = res2.pop("_ws")
    U = res2["U"]
    off = np.log(res2["kwargs_glm"]["size_factor"])
    df, est = cb.LFC(Y, np.c_[X, U], pd.DataFrame(A), np.c_[X_A, U], family=w["family"], offset=off,
                     usevar=w["usevar"], _glm=ws.glm, _trusted=True)
    if resident_ws is None:
        ws.close()
    n, p = Y.shape
    alt_iters = (len(res1["hist"]) - 1) + (len(res2["hist"]) - 1)
    irls = cglm.COUNTERS["irls_gene_iters"]
    stats = dict(alt_iters=alt_iters, irls_gene_iters=irls, updates=n * p * alt_iters + n * irls,
                 cell_updates=res1["total_cell_updates"] + res2["total_cell_updates"],
                 gene_rows=(alt_iters) * p,
                 cell_evals=res1["line_search_evals"][0] + res2["line_search_evals"][0],
                 gene_evals=res1["line_search_evals"][1] + res2["line_search_evals"][1],
                 gene_evals_onehot=res1["line_search_evals_onehot"][0] + res2["line_search_evals_onehot"][0],
                 onehot_cells=max(res1["line_search_evals_onehot"][1], res2["line_search_evals_onehot"][1]),
                 loop_seconds=res1["loop_seconds"] + res2["loop_seconds"], n=n, p=p,
                 gram_flops=cglm.COUNTERS["gram_flops"], gram_flops_structured=cglm.COUNTERS["gram_flops_structured"],
                 n_sig=int(np.nansum(df["padj"].to_numpy() < 0.05)))
    return df, stats


FULL_CALL = dict(n_ctrl=2000, n_perts=200, cells_per_pert=390, batch_size=15, max_cells=2000)


def full_call(rank, world, barrier):
    """The north_star's quoted call: ONE public ``gcate_lfc_batch(Y, X, A, 10, W_A=X_A, batch_size=15,
    max_cells=2000, n_ctrl=2000, family='nb', ...)`` on the 200-perturbation screen exactly as
    SURVEY.md section 8d (C3) spells it -- control subsampling, 14 batches, concatenated frame --
    wall-clock, host arrays in, DataFrame out, batches sharded over the ranks."""
    import causarray_b200 as cb
    from causarray_b200.synth import simulate_screen_fast
    from causarray_b200.prep import prep_causarray_data
    w, f = WORKLOAD, FULL_CALL
    t0 = time.perf_counter()
    sim = simulate_screen_fast(n_ctrl=f["n_ctrl"], n_perts=f["n_perts"], cells_per_pert=f["cells_per_pert"], p=w["p"],
                               r=w["r"], family=w["family"], seed=7, base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    Yc, _, X, X_A = prep_causarray_data(Y, A)
    Yc = np.ascontiguousarray(Yc.astype(w["counts_dtype"]))
    t_gen = time.perf_counter() - t0
    es = dict(rel_tol=w["es_rel_tol"], max_iters=w["es_max_iters"])
    kw = dict(W_A=X_A, batch_size=f["batch_size"], max_cells=f["max_cells"], n_ctrl=f["n_ctrl"], family=w["family"],
              lfc_kwargs=dict(usevar=w["usevar"]),
              gcate_kwargs=dict(disp_glm=sim["disp"], kwargs_es_1=dict(es), kwargs_es_2=dict(es)), random_state=0)
    from causarray_b200 import _timing
    # untimed warm-up of this call path: two batches (the timed steps above warmed the kernels, not the batch
    # driver: its first batch pays first-use costs of the row-gather upload and of the frame assembly)
    cb.gcate_lfc_batch(Yc, X, A[:, :2 * f["batch_size"]], w["r"], **kw)
    barrier()
    _timing.enable(True)
    _timing.reset()
    t0 = time.perf_counter()
    df = cb.gcate_lfc_batch(Yc, X, A, w["r"], **kw)
    barrier()
    dt = time.perf_counter() - t0
    phases = {k: round(v, 4) for k, v in _timing.TOTALS.items()}
    _timing.enable(False)
    n_batches = -(-f["n_perts"] // f["batch_size"])
    out = {"seconds": round(dt, 3), "phases_s_rank0": phases,
           "outside_phases_s_rank0": round(dt - sum(phases.values()), 4), "n_cells_total": int(Y.shape[0]), "n_genes": int(Y.shape[1]),
           "perturbations": f["n_perts"], "batches": n_batches, "n_gpus": world,
           "call": "gcate_lfc_batch(Y, X, A, 10, W_A=X_A, batch_size=15, max_cells=2000, n_ctrl=2000, family='nb', "
                   "lfc_kwargs=dict(usevar='unequal'), gcate_kwargs=dict(disp_glm=..., kwargs_es_1/2=dict(rel_tol=2e-4, "
                   "max_iters=30)), random_state=0)",
           "h2d_bytes": int(n_batches * (f["n_ctrl"] + f["batch_size"] * min(f["cells_per_pert"], f["max_cells"] // f["batch_size"]))
                            * Y.shape[1] * Yc.dtype.itemsize),
           "data_generation_s_untimed": round(t_gen, 1)}
    if rank == 0 and df is not None:
        out["result_rows"] = int(df.shape[0])
        out["n_significant"] = int(np.nansum(df["padj"].to_numpy() < 0.05))
    ref = full_size_reference_record()
    if ref is not None:
        cpu_s = ref["seconds_per_batch"] * n_batches
        out["cpu_reference_extrapolated_s"] = round(cpu_s, 0)
        out["cpu_reference_note"] = ("%d batches x %.1f s: the unmodified reference timed on ONE full-size batch "
                                     "(%d cells x %d genes) on %d host cores of a B200 box, %s; batches are independent "
                                     "and equally sized" % (n_batches, ref["seconds_per_batch"], ref["n_cells"],
                                                           ref["n_genes"], ref["cores"], ref["source"]))
        out["speedup_vs_cpu_reference"] = round(cpu_s / dt, 1)
    return out


def gene_shard_record(rank, world, barrier):
    """Gene-sharded alternating minimisation on the driver's record (north_star: "partitioned across
    the 8xB200 box by gene shards ... the cell-factor update all-reduces the per-cell gradient"):
    ONE problem -- the benchmark batch with four times the cells (15,980 x 8,000, k = 26) -- with its
    genes split over the N ranks, stage-1 iterations of ``update_with_mask`` (causarray/gcate_opt.py:
    148-206): the gene step is local, the cell step all-reduces the (n x r) gradient block, the n
    per-cell log-likelihoods and the n candidate objectives of every line-search round over NVLink.
    Strong scaling: the N = 1 line of the same bench run at --gpus 1 is the baseline."""
    import scipy.linalg
    import torch
    import torch.distributed as dist
    from causarray_b200.device import GcateDevice, comm_unique_id
    from causarray_b200.synth import simulate_screen_fast
    w = WORKLOAD
    sim = simulate_screen_fast(n_ctrl=4 * w["n_ctrl"], n_perts=w["n_perts"], cells_per_pert=4 * w["cells_per_pert"],
                               p=w["p"], r=w["r"], family=w["family"], seed=11, base_lo=-1.5, base_hi=1.5)
    Y, A = sim["Y"], sim["A"]
    n, p = Y.shape
    Xf = np.c_[np.ones((n, 1)), A]
    d, r = Xf.shape[1], w["r"]
    rng = np.random.default_rng(0)
    A0 = np.c_[Xf, rng.standard_normal((n, r)) * 0.05]
    B0 = np.c_[np.log(Y.mean(0) + 0.05)[:, None], rng.standard_normal((p, d - 1 + r)) * 0.05]
    Q1 = scipy.linalg.qr(Xf, mode="economic")[0]
    ls = np.log(np.maximum(Y.sum(1), 1))
    sf = np.exp(ls - ls.mean())
    edges = np.linspace(0, p, world + 1).astype(int)
    lo, hi = int(edges[rank]), int(edges[rank + 1])
    dev = GcateDevice(n, hi - lo, d, r, w["family"], 100.0)
    if world > 1:
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        dev.set_shard(rank, world, p, box[0])
    dev.load_counts(np.ascontiguousarray(Y[:, lo:hi], dtype=np.float64), sf, np.ascontiguousarray(sim["disp"][lo:hi]))
    dev.set_factors(A0, np.ascontiguousarray(B0[lo:hi]))
    dev.set_stage(P1=Q1, P2=None, lam=None, a=d)
    for _ in range(3):
        f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
    barrier()
    iters = 20
    t0 = time.perf_counter()
    for _ in range(iters):
        f = dev.update(1e3, 0.1, 0.5, 20, 1e-4)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    barrier()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ldg = (r + 3) // 4 * 4
    return {"workload": "alternating iteration (stage 1) of ONE fit, genes split over the ranks",
            "n_cells": int(n), "n_genes": int(p), "k": int(d + r), "n_gpus": world, "genes_per_rank": int(hi - lo),
            "ms_per_alt_iteration": round(1e3 * dt, 4), "cell_gene_updates_per_s": n * p / dt, "scaling": "strong",
            "objective_after_23_iterations": float(f),
            "allreduce_doubles_per_iteration": {"cell_gradient_and_loglik": int(n * (ldg + 1)),
                                                "line_search_objectives_per_round": int(n), "scalars": 2},
            "transport": {0: "single GPU", 1: "NCCL all-reduce",
                          2: "k_peer_allreduce over NVLink peer memory"}[int(dev.lib.ca_gcate_transport(dev.h))],
            "transport_rule": "peer memory on 2 and 4 ranks (8 with CA_PEER_REDUCE=1), NCCL otherwise"}


def cpu_step(batch, n_jobs):
    from oracle import pipeline as op
    w = WORKLOAD
    es = dict(rel_tol=w["es_rel_tol"], max_iters=w["es_max_iters"])
    g, l, st = op.gcate_lfc_one_batch_cpu(np.asarray(batch["Y"], dtype=np.float64), batch["X"], batch["A"], batch["X_A"], w["r"],
                                          family=w["family"], disp=batch["disp"], es1=dict(es), es2=dict(es),
                                          usevar=w["usevar"], n_jobs=n_jobs)
    return st


def reference_step(batch):
    """One batch through the REAL reference (staged oracle/_ref: numba alternating minimisation, the
    reference's own fit_gcate / LFC; per-gene GLM through the statsmodels stand-in)."""
    from oracle import refdrive
    w = WORKLOAD
    es = dict(rel_tol=w["es_rel_tol"], max_iters=w["es_max_iters"])
    res1, res2, df, est, st = refdrive.reference_fit_gcate_lfc(batch["Y"], batch["X"], batch["A"], batch["X_A"], w["r"],
                                                               family=w["family"], disp=batch["disp"], es1=dict(es),
                                                               es2=dict(es), usevar=w["usevar"])
    return st


def reference_arm(steps, warmup, p_s, cells_frac, kind):
    """Times the CPU implementation of the step on this box's host cores on a bounded sample of the
    same batch shape (fewer genes, and -- when the budget per step is small -- a quarter of the
    cells: both the per-cell and the per-gene loops of the reference carry a fixed cost per row, so a
    sample that keeps both dimensions in the hundreds stays close to the full-size throughput).
    kind = "reference": the unmodified reference package (numba JIT warmed up first, excluded from the
    timing); kind = "port": the C/OpenMP + numpy restatement of oracle/."""
    cores = os.cpu_count() or 1
    w = WORKLOAD
    jit_s = None
    if kind == "reference":
        os.environ.setdefault("NUMBA_NUM_THREADS", str(cores))
        from oracle import refdrive
        jit_s = refdrive.warm_up_jit(w["family"])
        step = reference_step
        how = ("unmodified reference package (oracle/_ref): numba update_with_mask on %d threads, fit_glm over joblib "
               "workers (n_jobs=-3) with the statsmodels stand-in of oracle/shims, its own LFC" % cores)
    else:
        from oracle import cport
        cport.set_threads(cores)
        step = lambda b: cpu_step(b, cores)
        how = ("C/OpenMP restatement of the alternating iteration (%d threads) + per-gene numpy IRLS over %d processes"
               % (cores, cores))
    times, upd = [], []
    stats = None
    n = None
    for s in range(warmup + steps):
        batch = make_batch(1000 + s, p=p_s, cells_frac=cells_frac)
        n = batch["Y"].shape[0]
        t0 = time.perf_counter()
        stats = step(batch)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
            upd.append(stats["cell_gene_updates"])
    T = sum(times)
    value = sum(upd) / T
    sample = ("one batch of the workload shape (r=%d, %d perturbations) restricted to %d of %d cells and %d of %d genes; %s"
              % (w["r"], w["n_perts"], n, w["n_ctrl"] + w["n_perts"] * w["cells_per_pert"], p_s, w["p"], how))
    return dict(value=value, ms_per_step=1e3 * T / max(1, len(times)), cores=cores, kind=kind, sample=sample,
                jit_warmup_s=jit_s, n_genes_sample=p_s, n_cells_sample=n,
                phases_s={k: round(v, 3) for k, v in stats.items() if k.startswith("t_")},
                work_last_step={k: int(stats[k]) for k in ("alt_iters", "irls_gene_iters") if k in stats})


def full_size_reference_record():
    """The reference timed at the FULL batch size on a B200 box's host cores (tools/ref_parity.py,
    committed record): context for the bounded samples above."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_config3_ref_parity.json")) as f:
            r = json.load(f)
        return {"source": "profiles/r02_config3_ref_parity.json (tools/ref_parity.py --config 3)",
                "n_cells": r["n"], "n_genes": r["p"], "seconds_per_batch": r["reference_seconds"],
                "cores": r["reference_cores"], "updates_per_s": r["reference_updates_per_s"],
                "tau_rel_max_vs_device": r["tau_rel_max"], "std_rel_max_vs_device": r["std_rel_max"],
                "rejections_differ": r["rejections_differ"]}
    except (OSError, ValueError, KeyError):
        return None


def sample_shape(kind, nsteps, genes=None, cells_frac=None):
    """Sample of the CPU arms: about 240 s of CPU work over all steps (reference: ~7e6 updates/s on 16
    cores, ~90 update-equivalents per cell and gene of a step)."""
    w = WORKLOAD
    if genes is not None:
        return int(genes), (1.0 if cells_frac is None else float(cells_frac))
    cores = os.cpu_count() or 1
    rate = (7.0e6 if kind == "reference" else 2.5e7) * cores / 16.0
    units = 240.0 / max(1, nsteps) * rate / 90.0
    n_full = w["n_ctrl"] + w["n_perts"] * w["cells_per_pert"]
    if cells_frac is None:
        cells_frac = 1.0 if units >= n_full * 1000 else 0.25
    n_s = n_full * cells_frac
    p_s = int(max(256, min(w["p"], units / n_s))) // 8 * 8
    return p_s, cells_frac


def reference_kind():
    from oracle import ref_shim
    return "reference" if ref_shim.reference_available() else "port"


def run_reference(args, rank, world):
    """CPU arm: rank 0 only; bounded sample = gene subsample of the same batch shape."""
    if rank != 0:
        return
    kind = args.cpu_impl or reference_kind()
    p_s, cf = sample_shape(kind, args.steps + args.warmup, args.cpu_genes, args.cpu_cells_frac)
    r = reference_arm(args.steps, args.warmup, p_s, cf, kind)
    w = WORKLOAD
    n = w["n_ctrl"] + w["n_perts"] * w["cells_per_pert"]
    out = {"impl": "reference", "metric": "gcate_lfc_cell_gene_updates_per_s", "value": r["value"],
           "unit": "cell*gene updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": w["name"], "n_cells": n, "n_cells_sample": r["n_cells_sample"],
                      "n_genes_sample": p_s, "n_genes": w["p"], "r": w["r"],
                      "perturbations": w["n_perts"], "family": w["family"]},
           "cpu_baseline": {"value": r["value"], "unit": "cell*gene updates/s", "cores": r["cores"], "kind": r["kind"],
                            "sample": r["sample"], "jit_warmup_s_excluded": r["jit_warmup_s"],
                            "full_size_record": full_size_reference_record()},
           "e2e": {"value": r["value"], "unit": "cell*gene updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "phases_s": r["phases_s"], "work_last_step": r["work_last_step"]}
    emit_json(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-genes", type=int, default=None, help="gene subsample of the CPU arms")
    ap.add_argument("--cpu-cells-frac", type=float, default=None, help="cell subsample of the CPU arms")
    ap.add_argument("--cpu-impl", default=None, choices=["reference", "port"],
                    help="CPU arm: the staged reference package (default when present) or the oracle port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-call", action="store_true", help="skip the whole-screen gcate_lfc_batch measurement")
    ap.add_argument("--no-gene-shard", action="store_true", help="skip the gene-sharded strong-scaling sub-record")
    ap.add_argument("--genes", type=int, default=None, help="override p (debug only; makes the number invalid)")
    args = ap.parse_args()
    capture_stdout()
    warnings.simplefilter("ignore")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    from causarray_b200 import _lib as L
    L.load()
    L.check(L._lib.ca_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from causarray_b200.optim import BatchWorkspace
    w = WORKLOAD
    nb = args.warmup + args.steps
    batches = [make_batch(100 * rank + s, p=args.genes) for s in range(min(nb, 4))]
    n, p = batches[0]["Y"].shape
    d = batches[0]["X"].shape[1] + batches[0]["A"].shape[1]

    # ---------------- resident arm: counts uploaded outside the timed region ----------------
    def resident_pass(nsteps, timed):
        tot_updates, stats_all = 0, []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        elapsed = 0.0
        for s in range(nsteps):
            b = batches[s % len(batches)]
            ws = BatchWorkspace.upload(b["Y"], w["family"], d, w["r"])
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            df, st = gpu_step(b, resident_ws=ws)
            ev1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            # device clock between the two events on the launching (default) stream; the step ends with
            # host work (p-values, BH, frame), which the second event -- recorded after it -- includes.
            # The host clock is the cross-check (they agree to ~0.1 ms); the larger one is reported.
            elapsed += max(ev0.elapsed_time(ev1) * 1e-3, wall)
            ws.close()
            tot_updates += st["updates"]
            stats_all.append(st)
        return elapsed, tot_updates, stats_all

    from causarray_b200 import _timing
    resident_pass(args.warmup, False)
    _timing.enable(True)
    _timing.reset()
    L.profile_enable(True)
    L.profile_reset()
    launches0 = L.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_res, upd_res, stats_res = resident_pass(args.steps, True)
    barrier()
    prof = L.profile_snapshot()
    launches = L.launch_count() - launches0
    L.profile_enable(False)
    phases = {k: round(1e3 * v / args.steps, 2) for k, v in _timing.TOTALS.items()}
    _timing.enable(False)

    # ---------------- end-to-end arm: host arrays in, DataFrame out ----------------
    for s in range(min(args.warmup, 1)):
        gpu_step(batches[s % len(batches)])
    barrier()
    t0 = time.perf_counter()
    upd_e2e = 0
    d2h = 0
    for s in range(args.steps):
        df, st = gpu_step(batches[s % len(batches)])
        upd_e2e += st["updates"]
        d2h = int(df.memory_usage(index=False).sum())
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    e2e_full = None
    if not args.no_full_call and args.genes is None:
        try:
            e2e_full = full_call(rank, world, barrier)
        except Exception as exc:  # noqa: BLE001  (never lose the main line over the extra measurement)
            e2e_full = {"error": "%s: %s" % (type(exc).__name__, exc)}
    gene_shard = None
    if not args.no_gene_shard and args.genes is None:
        try:
            gene_shard = gene_shard_record(rank, world, barrier)
        except Exception as exc:  # noqa: BLE001
            gene_shard = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # max over ranks of the times, sum over ranks of the work
    if world > 1:
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([float(upd_res), float(upd_e2e)], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        t_res, t_e2e = t.tolist()
        upd_res, upd_e2e = u.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    # per-kernel-class algorithmic bytes over the timed resident steps (rank 0)
    st_sum = {k: sum(s[k] for s in stats_res) for k in ("cell_updates", "gene_rows", "cell_evals", "gene_evals",
                                                         "irls_gene_iters", "alt_iters", "gram_flops",
                                                         "gram_flops_structured", "gene_evals_onehot")}
    n_tr = max(s["onehot_cells"] for s in stats_res)
    alg_bytes = {
        # Y read once per active row: cells x p + genes x n doubles
        "rowpass_grad": 8.0 * (st_sum["cell_updates"] * p + st_sum["gene_rows"] * n),
        # one read of the row of Y per candidate evaluation
        # (stage-2 gene candidates of the one-hot search read the counts and the linear predictor of the
        # TREATED cells only: 16 B per entry)
        "search": 8.0 * (st_sum["cell_evals"] * p + (st_sum["gene_evals"] - st_sum["gene_evals_onehot"]) * n),
        "search_onehot": 16.0 * st_sum["gene_evals_onehot"] * n_tr,
        # one read of the gene's counts per IRLS iteration
        "glm_pass": 8.0 * st_sum["irls_gene_iters"] * n,
    }
    kern = {}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, cnt) in prof.items():
        kern[name] = {"ms": round(ms, 3), "launches": cnt, "share": round(ms / tot_ms, 4)}
        if name in alg_bytes and ms > 0:
            kern[name]["achieved_gbs"] = round(alg_bytes[name] / (ms * 1e-3) / 1e9, 1)
    # FP64 peak: DMMA (mma.sync m8n8k4.f64) and DFMA measured on this pool's B200 with tools/fp64_microbench
    # (both 36.6-37.1 TFLOP/s: the FP64 tensor path and the FP64 ALU share one pipe at one rate);
    # MEASURED_PEAKS.json holds no FP64 figure and tcgen05 has no FP64 kind
    fp64_peak_tflops, fp64_src = 36.5, "hard-coded round-1 measurement"
    try:
        with open(os.path.join(ROOT, "profiles", "r02_fp64_microbench.json")) as f:
            mb = json.load(f)
        fp64_peak_tflops = float(min(mb["dfma_tflops"], mb["dmma_m8n8k4_tflops"]))
        fp64_src = "profiles/r02_fp64_microbench.json (tools/fp64_microbench.cu: min of DFMA and DMMA m8n8k4)"
    except (OSError, ValueError, KeyError):
        pass
    k_tot = d + w["r"]
    # Algorithmic FP64 work per matrix entry.  A likelihood evaluation needs the
    # linear predictor (2 k flop), one exp, one log1p and the NB / Poisson arithmetic: 27 FP64
    # instructions = 54 flop-equivalents in the leanest straight-line form we found; a gradient pass adds
    # the reciprocal of the NB weight (6 instructions) and the contraction of dl with the r (or a)
    # moving columns (2 x 16 flop).
    flops_eval = 2.0 * k_tot + 54.0
    flops_grad = 2.0 * k_tot + 54.0 + 12.0 + 32.0
    alg_flops = {
        # one-hot stage-2 candidates: ~19 FP64 instructions per treated entry (no matrix product, no exp)
        "search": flops_eval * (st_sum["cell_evals"] * p + (st_sum["gene_evals"] - st_sum["gene_evals_onehot"]) * n),
        "search_onehot": 38.0 * st_sum["gene_evals_onehot"] * n_tr,
        "rowpass_grad": flops_grad * (st_sum["cell_updates"] * p + st_sum["gene_rows"] * n),
        # IRLS: the round-1 formula 2 n [kx(kx+1)/2 + 2 kx] per gene-iteration, kept so that the
        # fraction is comparable between rounds; the structure-aware count is reported beside it
        "glm_pass": st_sum["gram_flops"],
    }
    if prof.get("glm_pass", (0, 0))[0] > 0:
        kern["glm_pass"]["achieved_tflops_fp64"] = round(st_sum["gram_flops"] / (prof["glm_pass"][0] * 1e-3) / 1e12, 2)
    dom = max((k for k in alg_bytes if prof.get(k, (0, 0))[0] > 0), key=lambda k: prof[k][0], default=None)
    # DRAM bytes of one captured launch per kernel class (ncu --set full, committed next to its summary)
    traffic = {}
    for fn in ("r01_traffic.json", "r02_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                traffic.update(json.load(f))
        except (OSError, ValueError):
            pass

    def kernel_roofline(name):
        ms, cnt = prof[name]
        tr = traffic.get(name)
        ach = alg_flops[name] / (ms * 1e-3) / 1e12
        r = {"kernel": name, "bound": "tensor", "achieved": round(ach, 2), "peak": round(fp64_peak_tflops, 2),
             "unit": "TFLOP/s", "frac": round(ach / fp64_peak_tflops, 4),
             "peak_source": "FP64 pipe (DMMA mma.sync m8n8k4.f64 = DFMA rate) measured on this pool's B200: " + fp64_src
                            + "; MEASURED_PEAKS.json holds only bf16 tensor and HBM figures, tcgen05 has no FP64 kind",
             "hbm_view": {"achieved_gbs": round(alg_bytes[name] / (ms * 1e-3) / 1e9, 1), "peak_gbs": hbm_peak,
                          "frac": round(alg_bytes[name] / (ms * 1e-3) / 1e9 / hbm_peak, 4),
                          "algorithmic_bytes": "8 B per cell*gene entry and pass (one read of Y)",
                          "peak_source": peak_src}}
        if name == "glm_pass":
            r["algorithmic_flops"] = ("2 n [kx(kx+1)/2 + 2 kx] per gene and IRLS iteration (round-1 / SURVEY 8d formula; "
                                      "it counts the products with the one-hot treatment block that the kernel no longer forms)")
            sa = st_sum.get("gram_flops_structured", 0.0)
            if sa:
                r["structure_aware"] = {"achieved": round(sa / (ms * 1e-3) / 1e12, 2),
                                        "frac": round(sa / (ms * 1e-3) / 1e12 / fp64_peak_tflops, 4),
                                        "algorithmic_flops": "2 n [kd(kd+1)/2 + 2 kd + 65] per gene-iteration, kd = dense "
                                                             "design columns (pair products, right-hand side, linear "
                                                             "predictor, ~65 flop-equivalents of exp / log / reciprocals)"}
        elif name == "search":
            r["algorithmic_flops"] = "(2 k + 54) per evaluated entry of a line-search candidate, k = %d" % k_tot
        elif name == "search_onehot":
            r["algorithmic_flops"] = ("38 per TREATED entry of a stage-2 gene candidate (%d of %d cells; the control cells "
                                      "and the matrix product drop out): the kernel is bound by latency and shared-memory "
                                      "traffic, not by the FP64 pipe" % (n_tr, n))
            r["hbm_view"]["algorithmic_bytes"] = "16 B per treated cell*gene entry and candidate (count + linear predictor)"
        else:
            r["algorithmic_flops"] = "(2 k + 98) per entry of a gradient pass, k = %d" % k_tot
        r["note"] = ("bound by the shared FP64 pipe (DMMA + DFMA + table-driven exp / log), not by HBM: the HBM view is "
                     "listed for reference; ncu pipe utilisation in profiles/")
        r["avg_launch_ms"] = round(ms / max(1, cnt), 4)
        r["launches"] = cnt
        r["traffic"] = tr["dram_bytes"] if tr else None
        if tr:
            r["traffic_note"] = "DRAM read+write bytes of one captured launch (%s); algorithmic bytes of that launch: %s" \
                                % (tr["capture"], tr.get("algorithmic_bytes") or tr.get("algorithmic_note"))
        return r

    roofline = None
    if dom is not None:
        roofline = kernel_roofline(dom)
        roofline["others"] = [kernel_roofline(k) for k in ("search", "glm_pass", "rowpass_grad", "search_onehot")
                              if k != dom and prof.get(k, (0, 0))[0] > 0]

    out = {"metric": "gcate_lfc_cell_gene_updates_per_s", "value": upd_res / t_res, "unit": "cell*gene updates/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": w["name"], "n_cells": n, "n_genes": p, "r": w["r"],
                      "perturbations_per_batch": w["n_perts"], "family": w["family"],
                      "batches_per_step_per_gpu": 1, "parallelism": "batch-sharded x%d" % world,
                      "counts_dtype_host": w["counts_dtype"],
                      "l2_policy": "inputs larger than L2 (Y and Y^T are 256 MB each at fp64)"},
           "e2e": {"value": upd_e2e / t_e2e, "unit": "cell*gene updates/s", "ms_per_step": 1e3 * t_e2e / args.steps,
                   "h2d_bytes_per_step": int(batches[0]["Y"].nbytes + batches[0]["X"].nbytes
                                             + batches[0]["A"].nbytes + batches[0]["X_A"].nbytes),
                   "d2h_bytes_per_step": d2h},
           "e2e_full": e2e_full, "gene_shard": gene_shard,
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kern,
           "phases_ms_per_step": phases,
           "work_per_step": {k: v / args.steps for k, v in st_sum.items()},
           "n_significant_last_step": stats_res[-1]["n_sig"]}

    if not args.no_cpu_baseline and world == 1:
        # CPU baseline in a child process (the reference brings numba / OpenMP pools of its own): the
        # real reference when it is staged, and the oracle port beside it as a second, labelled number
        def child(kind, genes, frac):
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-impl", kind, "--steps", "1",
                   "--warmup", "0", "--cpu-genes", str(genes), "--cpu-cells-frac", str(frac)]
            try:
                res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=900)
                line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
                j = json.loads(line)
                cb = j["cpu_baseline"]
                cb["seconds"] = round(j["ms_per_step"] * 1e-3, 2)
                cb["phases_s"] = j.get("phases_s")
                return cb
            except Exception as e:  # noqa: BLE001
                return {"error": "%s: %s" % (type(e).__name__, e)}
        kind = reference_kind()
        # ~25 s of CPU work: a quarter of the cells x 2000 genes for the reference, all cells x 400 genes for the port
        out["cpu_baseline"] = (child(kind, args.cpu_genes or 2000, args.cpu_cells_frac or 0.25) if kind == "reference"
                               else child(kind, args.cpu_genes or 400, args.cpu_cells_frac or 1.0))
        if kind == "reference":
            out["cpu_baseline_port"] = child("port", args.cpu_genes or 400, args.cpu_cells_frac or 1.0)
    emit_json(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
