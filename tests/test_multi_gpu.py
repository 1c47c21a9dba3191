"""GPU, needs >= 2 devices (skipped on a single-GPU box): the torchrun parity script as a pytest case -- source-sharded
lnprob over NCCL and over the peer-memory exchange equals the single-GPU engine and the oracle, and the multi-GPU
device-resident sampler reproduces its host replay (tests/run_multi_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least two GPUs")
def test_source_sharded_parity_under_torchrun():
    n = min(_ngpu(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', '29577', os.path.join(ROOT, 'tests', 'run_multi_gpu_parity.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'MULTI_GPU_PARITY OK' in out.stdout, out.stdout[-2000:]
