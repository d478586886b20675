"""cProfile of one benchmark step sorted by own time (host-side glue; not a benchmark)."""
import cProfile, io, os, pstats, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import bench
b = bench.make_batch(7)
bench.gpu_step(b)
bench.gpu_step(b)
pr = cProfile.Profile()
pr.enable()
bench.gpu_step(b)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
