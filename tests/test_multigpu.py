"""Multi-GPU tests (need >= 2 CUDA devices): fused outer step over peer memory vs NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("chunks", [1, 4])
def test_fused_outer_step_matches_nccl_reference(chunks):
    """chunks=1: plain A|B|C schedule; chunks=4: software-pipelined schedule with role-split CTAs."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_outer_check.py"), "--numel",
           str(16 * 1024 * 1024), "--chunks", str(chunks)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=550, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"correct": true' in r.stdout


@pytest.mark.timeout(600)
def test_nvls_multimem_outer_step_matches_reference():
    """Same check with the buffers bound to an NVSwitch multicast object (multimem.ld_reduce/st)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tests", "mgpu_outer_check.py"), "--numel",
           str(16 * 1024 * 1024), "--nvls", "--chunks", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=550, cwd=ROOT)
    if r.returncode != 0 and "NvlsUnavailable" in r.stderr:
        pytest.skip("NVSwitch multicast could not be set up on this box: " + r.stderr[-300:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"correct": true' in r.stdout


@pytest.mark.timeout(600)
def test_kill_and_rejoin_over_peer_memory():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(_ngpu(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "mgpu_fault_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=550, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"correct": true' in r.stdout
