"""Multi-GPU parity (one process per GPU, NCCL transport) — runs tools/multi_check.py under torchrun when the box
has at least two GPUs; the single-GPU boxes of the round-end run skip it (the multi-rank path is covered there by
the thread transport in test_gpu_steps.py / test_gpu_hybrid.py and on the CPU by test_multirank_host.py; `bench.py --gpus N`
runs the same check over NCCL ahead of its timed region)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("replica,hybrid", [(1, 1), (1, 0), (0, 0)])       # hybrid solve (default), replicated, distributed
@pytest.mark.parametrize("kind", ["warm", "cold"])
def test_two_gpus_nccl(kind, replica, hybrid):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_check.py"), kind, "4", str(replica), str(hybrid)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("] ok: ") == 2
