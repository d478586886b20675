"""cProfile of the whole-screen gcate_lfc_batch call of bench.full_call (host-side overheads)."""
import cProfile, io, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import bench
import causarray_b200 as cb
from causarray_b200 import _timing
from causarray_b200.synth import simulate_screen_fast
from causarray_b200.prep import prep_causarray_data

w, f = bench.WORKLOAD, bench.FULL_CALL
sim = simulate_screen_fast(n_ctrl=f["n_ctrl"], n_perts=f["n_perts"], cells_per_pert=f["cells_per_pert"], p=w["p"], r=w["r"],
                           family=w["family"], seed=7, base_lo=-1.5, base_hi=1.5)
Y, A = sim["Y"], sim["A"]
t0 = time.perf_counter()
Yc, _, X, X_A = prep_causarray_data(Y, A)
Yc = np.ascontiguousarray(Yc.astype(w["counts_dtype"]))
print("prep_causarray_data + astype: %.2f s" % (time.perf_counter() - t0))
es = dict(rel_tol=w["es_rel_tol"], max_iters=w["es_max_iters"])
kw = dict(W_A=X_A, batch_size=f["batch_size"], max_cells=f["max_cells"], n_ctrl=f["n_ctrl"], family=w["family"],
          lfc_kwargs=dict(usevar=w["usevar"]), gcate_kwargs=dict(disp_glm=sim["disp"], kwargs_es_1=dict(es), kwargs_es_2=dict(es)),
          random_state=0)
cb.gcate_lfc_batch(Yc[:, :], X, A[:, :30], w["r"], **{**kw, "n_batches": None})   # warm-up: 2 batches
_timing.enable(True); _timing.reset()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
df = cb.gcate_lfc_batch(Yc, X, A, w["r"], **kw)
pr.disable()
print("full call: %.2f s" % (time.perf_counter() - t0))
print({k: round(v, 3) for k, v in _timing.TOTALS.items()}, "sum", round(sum(_timing.TOTALS.values()), 3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:6000])
