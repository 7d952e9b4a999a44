"""Multi-GPU parity (needs >= 2 GPUs on the box): the sharded step under torchrun
equals the oracle's R-replica simulation, and every rank ends with identical weights."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('args', [['fp32', 'simt', 'global'], ['fp32', 'simt', 'local'], ['bf16', 'tc', 'global', 'graph'],
                                  ['fp32', 'simt', 'global', 'nccl']],
                         ids=['fp32-syncbn-peer', 'fp32-localbn-peer', 'bf16-peer-graph', 'fp32-syncbn-nccl'])
def test_two_rank_step_matches_oracle(args):
    """Default: SyncBN statistics and the embedding / lse all-gathers through the NVLink peer-memory kernels
    (csrc/comm.cu); 'nccl': the same collectives through torch.distributed.  'graph': additionally two segmented
    CUDA-graph replays against two eager steps."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ)
    if args[-1] == 'nccl':
        args = args[:-1]
        env['SIMCLR_COMM'] = 'nccl'
    else:
        env['SIMCLR_COMM'] = 'peer_required'
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    assert 'MULTI_GPU_CHECK' in out, out[-3000:]
    line = [l for l in out.splitlines() if l.startswith('MULTI_GPU_CHECK')][0]
    print(line)
    assert '"ok": true' in line, line
    assert 'MULTI_GPU_WEIGHTS_IDENTICAL 1' in out
    assert ('"collectives": "nccl"' in line) == (env['SIMCLR_COMM'] == 'nccl'), line
    if 'graph' in args:
        gl = [l for l in out.splitlines() if l.startswith('MULTI_GPU_GRAPH')]
        assert gl and '"ok": true' in gl[0], out[-3000:]
        print(gl[0])
    assert r.returncode == 0, out[-2000:]
