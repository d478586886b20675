"""Is host-side threading slower under torchrun / after NCCL init?  (not a benchmark)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from causarray_b200 import inference as inf

rng = np.random.default_rng(0)
t = rng.standard_normal((15, 8000)) * 2
df = rng.uniform(50, 400, (15, 8000))


def T(f, n=7):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return "min %.2f med %.2f ms" % (min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3)


rank = int(os.environ.get("RANK", "0"))
print(rank, "OMP", os.environ.get("OMP_NUM_THREADS"), "threads", inf._N_THREADS, "affinity", len(os.sched_getaffinity(0)),
      "before init:", T(lambda: inf._tail_parallel(t, df)), flush=True)
if "--nccl" in sys.argv:
    import torch, torch.distributed as dist
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dist.barrier()
    torch.cuda.synchronize()
    print(rank, "after NCCL init: affinity", len(os.sched_getaffinity(0)), T(lambda: inf._tail_parallel(t, df)), flush=True)
    x = torch.zeros(4, device="cuda")
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    print(rank, "after all_reduce:", T(lambda: inf._tail_parallel(t, df)), flush=True)
    dist.destroy_process_group()
