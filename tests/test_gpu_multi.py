"""Multi-GPU PLR equivalence (NCCL): env-sharded ranks with the level-encoding / episode-record / done-flag all-gathers of
dcd_isaac_b200.distributed.ShardedPLR reproduce the single-GPU sampler, level store and rollout tensors bit for bit
(tests/dist_plr_worker.py does the comparison on rank 0).  Needs >= 2 GPUs (`gpurun --gpus 2`); the single-GPU variant runs
the same worker unsharded so that the code path is exercised on a 1-GPU box too."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

WORKER = os.path.join(ROOT, 'tests', 'dist_plr_worker.py')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plr_cycles_single_gpu_worker():
    out = subprocess.run([sys.executable, WORKER, '--envs', '32', '--T', '40', '--cycles', '4', '--buffer', '48'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert 'SINGLE ok' in out.stdout


@pytest.mark.parametrize('strategy', ['positive_value_loss', 'grounded_signed_value_loss'])
def test_sharded_plr_matches_single_gpu(strategy):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    world = 2 if torch.cuda.device_count() < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), WORKER, '--envs', '64', '--T', '48', '--cycles', '6', '--strategy', strategy]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2500:] + out.stderr[-2500:]
    assert 'SHARDED_PLR MATCH' in out.stdout, out.stdout[-2500:]
