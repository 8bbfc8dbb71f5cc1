"""Multi-GPU parity on real devices: one process per GPU under torchrun (NCCL + CUDA-IPC peer memory).
Skipped below 2 GPUs; the same host logic runs on the CPU with gloo in tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _torchrun(nproc, script, *args, port=29531):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_sobol_iteration_matches_single_gpu(nproc):
    """pool / nfev identical for the fused, two-phase p2p and all_gather exchanges and for one GPU"""
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = _torchrun(nproc, "check_sharded.py", "22", port=29531 + nproc)
    assert r.returncode == 0 and ": OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nproc", [2, 4])
def test_sharded_lattice_generation_matches_single_gpu(nproc):
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = _torchrun(nproc, "check_sharded_lattice.py", port=29541 + nproc)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nproc", [2, 4])
def test_sharded_host_buffer_entry_points_match_single_gpu(nproc):
    """the C-ABI host-buffer path, one ctx per rank, against shgo_b200_iterate_sobol_host on one GPU"""
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = _torchrun(nproc, "check_sharded_host.py", "20", port=29551 + nproc)
    assert r.returncode == 0 and ": OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
