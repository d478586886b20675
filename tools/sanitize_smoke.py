"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck, one tool
per call): the smoke checks of __graft_entry__ plus one tiny fit_gcate + LFC through the public API, so
that every kernel class (row passes, line search, persistent IRLS with one-hot groups and the lasso
refit, AIPW, inference tail, size factors, CSR / typed uploads) launches at least once.

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")

import __graft_entry__ as ge  # noqa: E402

ge.smoke()
import causarray_b200 as cb  # noqa: E402
from causarray_b200.synth import simulate_screen  # noqa: E402
from causarray_b200.prep import prep_causarray_data  # noqa: E402

sim = simulate_screen(n_ctrl=90, n_perts=3, cells_per_pert=25, p=120, r=2, d_cov=1, family="nb", seed=4,
                      base_lo=0.0, base_hi=2.0)
Y, A = sim["Y"], sim["A"]
_, _, X, X_A = prep_causarray_data(Y, A)
es = dict(rel_tol=2e-4, max_iters=6)
r1, r2 = cb.fit_gcate(Y.astype(np.int32), X, A, 2, family="nb", disp_glm=sim["disp"], offset=True, kwargs_es_1=dict(es),
                      kwargs_es_2=dict(es))
ws = r2.pop("_ws")
off = np.log(r2["kwargs_glm"]["size_factor"])
df, _ = cb.LFC(Y, np.c_[X, r2["U"]], pd.DataFrame(A), np.c_[X_A, r2["U"]], family="nb", offset=off, usevar="unequal",
               _glm=ws.glm, _trusted=True)
ws.close()
print("sanitize_smoke ok: %d rows, %d significant" % (df.shape[0], int(np.nansum(df["padj"] < 0.05))))
