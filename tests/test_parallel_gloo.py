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
