"""Gene-sharded alternating minimisation over two GPUs (NCCL all-reduces of the cell-step sums)
against the real-reference golden.  Needs two devices; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("fam", ["nb", "poisson"])
def test_gene_sharded_alter_min_two_gpus(fam):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "mgpu", "sharded_altmin.py"), fam]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    print(out.stderr[-2000:])
    assert out.returncode == 0


def test_gene_sharded_fit_gcate_two_gpus():
    """fit_gcate(gene_shard=True) end to end (GLM initialisation, sharded SVD, both stages) against the
    same fit on one GPU."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(root, "tests", "mgpu", "sharded_fit_gcate.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    print(out.stderr[-2000:])
    assert out.returncode == 0
