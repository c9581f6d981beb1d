"""Multi-GPU parity: distributed assembly + Krylov solve (NCCL halo exchange + all-reduce) against the single-GPU solve.
Needs >= 2 GPUs (skipped otherwise); the world_size-2 host logic is covered on CPU by tests/test_partition_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distributed_solve_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "dist_solve_check.py"), "12"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "DIST_CHECK_PASS" in r.stdout
