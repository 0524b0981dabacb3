"""N>1: one process per GPU, NCCL send/recv halo exchange, against the oracle's emulated MPI decomposition."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("layout", [(2, 1), (1, 2), (2, 2)])
def test_nccl_halo_exchange_matches_oracle(tmp_path, layout, p2p):
    """p2p = 1: the peer-to-peer exchange (halo_push / halo_wait / halo_pull over CUDA IPC memory); 0: pack + ncclSend/Recv + unpack"""
    n = layout[0] * layout[1]
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mgpu_worker.py"), str(tmp_path), str(layout[0]), str(layout[1]), "30", str(p2p)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count(" ok:") == n
    assert p.stdout.count(f" p2p={p2p}") == n, p.stdout[-2000:]          # the path that was asked for is the one that ran


@pytest.mark.parametrize("n", [2, 3])
def test_psv_nccl(tmp_path, n):
    """swpc_psv over NCCL: column exchange (m_global.f90:296-420 of swpc_psv), max-amplitude reduce and the snapshot
    sum-reduce onto the I/O ranks, against the oracle's emulated ranks."""
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    port = 29300 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mgpu_psv_worker.py"), str(tmp_path), "40"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count(" psv ok:") == n


@pytest.mark.parametrize("p2p", [1, 0])
def test_a_silent_neighbour_is_detected(tmp_path, p2p):
    """Failure detection (the reference aborts cleanly, m_report.f90:144-151): rank 1 stops taking part after 12
    steps; rank 0's time loop must return an error within the communication time-out (4 s here) -- through the bounded spin of
    halo_wait (peer-to-peer transport) or the stream poll + ncclCommAbort (NCCL transport) -- instead of hanging."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29000 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mgpu_fail_worker.py"), str(tmp_path), str(p2p)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=90)
    assert "rank 0 aborted cleanly" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
    assert "rank 1 left the loop" in p.stdout
