"""-m gpu, real multi-GPU parity: one process per GPU under torchrun (spawned here when the box has >= 2 GPUs),
every rank owning its Ndw shard exactly as direct_mpi shards it (ED_HAMILTONIAN.f90:92-105).  H x v through all
three transpose back-ends (copy-engine exchange over CUDA-IPC windows, peer-memory stores, NCCL send/recv), the
Krylov coefficients with complex and real start vectors and a shard-local c^+ are compared on rank 0 with the
oracle's simulated-MPI path (tools/spmd_check.py does the work; this file only launches it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_spmd_parity_under_torchrun():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (the -m gpu tests with simulated ranks cover the sharded code path on one)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "spmd_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout[-3000:] + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "PASS" in r.stdout and "FAIL" not in r.stdout, tail
