"""Multi-GPU (domain decomposition) parity: launches tests/mgpu_check.py under torchrun with one rank
per GPU when the box has at least two. On the single-GPU box this is skipped; the host-side logic is
covered by tests/test_domain_host.py (gloo, CPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_domain_decomposition(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc),
           os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mgpu ok" in r.stdout


@pytest.mark.parametrize("nproc", [2, 8])
def test_peer_memory_mailboxes(nproc):
    """The per-step one-word reductions over NVLink peer memory (CUDA IPC mailboxes) with many ranks."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29520 + nproc),
           os.path.join(ROOT, "tests", "p2p8_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "p2p8 ok" in r.stdout and "p2p_mailboxes=1" in r.stdout
