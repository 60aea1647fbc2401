"""Multi-GPU (NCCL) parity of the slab-decomposed operator and CG against the CPU oracle; needs
>= 2 GPUs on the box (skipped otherwise).  Runs tests/slab_check.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def test_slab_cg_matches_oracle_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29541', os.path.join(ROOT, 'tests', 'slab_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]+out.stderr[-2000:]
    assert 'FAIL' not in out.stdout and out.stdout.count(' ok') >= 3


def test_slab_pipeline_single_rank():
    """the chunked zero-copy exchange pipeline and the device-scalar CG on one rank (the exchange is a
    copy): every kernel of the multi-GPU path runs on a 1-GPU box"""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '1',
           '--master-addr', '127.0.0.1', '--master-port', '29543', os.path.join(ROOT, 'tests', 'slab_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:]+out.stderr[-2000:]
    assert 'FAIL' not in out.stdout and out.stdout.count(' ok') >= 12
