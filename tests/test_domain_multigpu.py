"""Slab decomposition across REAL GPUs (needs >= 2 devices; skipped otherwise):
NVLink peer-memory halo exchange, CUDA-graph step and the e2e call, each rank
checked against a single-GPU evaluation (tools/dd_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs two GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_slabs(peer):
    env = dict(os.environ, TAB_DD_PEER=peer)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port',
           '29541' if peer == '1' else '29542', os.path.join(ROOT, 'tools', 'dd_check.py'),
           '8']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'FAIL' not in res.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs two GPUs")
def test_two_rank_atomic_nn_slabs():
    """AtomicNN decomposition (2 rc halo, NCCL send/recv ring; tools/dd_atomic_check.py)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29543',
           os.path.join(ROOT, 'tools', 'dd_atomic_check.py'), '12', '3']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'FAIL' not in res.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs two GPUs")
def test_two_rank_structure_parallel_training():
    """Config 4: per-rank sub-batches, one flat NCCL gradient all-reduce
    (tools/train_check.py)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29544',
           os.path.join(ROOT, 'tools', 'train_check.py'), '4']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'FAIL' not in res.stdout
