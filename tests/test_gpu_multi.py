"""NCCL path of the Newton linear solve on >= 2 GPUs of one node (skipped on a single-GPU box): halo exchange, owned-row
dot products + all-reduce, distributed GMRES inside the Newton loop; see scripts/multi_gpu_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,dim,size,partition", [(2, 2, 6, "structured"), (2, 3, 3, "structured"), (2, 2, 6, "rcb"),
                                                     (2, 3, 3, "rcb")])
def test_distributed_newton_matches_single_gpu(world, dim, size, partition):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29617 + dim + (10 if partition == "rcb" else 0)),
           os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--size", str(size), "--dim", str(dim), "--partition", partition]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
