"""Data-parallel fused step on 2 GPUs of one box (skipped on a single-GPU box): NCCL gradient sum in GEMM order overlapped with the
backward pass, 1/world inside the fused Adam, early optimiser group behind the collective -- the replicas must stay bit-identical."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_fused_step_keeps_replicas_identical():
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), os.path.join(ROOT, 'tools', 'dp_check.py')], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    res = json.loads(line)
    print(res)
    assert res['replicas_identical'] and res['losses_finite']
