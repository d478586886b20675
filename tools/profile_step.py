"""One short bench-shaped step for ncu / cProfile (not a benchmark)."""
import cProfile
import io
import os
import pstats
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import bench  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
bench.WORKLOAD["es_max_iters"] = iters
b = bench.make_batch(7)
if mode == "cprofile":
    bench.gpu_step(b)
    pr = cProfile.Profile()
    pr.enable()
    t0 = time.perf_counter()
    df, st = bench.gpu_step(b)
    dt = time.perf_counter() - t0
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())
    print("step seconds", dt, st)
else:
    df, st = bench.gpu_step(b)
    print(st)
