"""Multi-rank tests: halo plan + exchange under gloo (CPU, world_size 2 and 3), and the decomposed
RK3 step against the single-GPU run under NCCL (needs >= 2 GPUs)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn_cpu(rank, world, port):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import dist_worker
    dist_worker.run_cpu(rank, world, port)


@pytest.mark.parametrize('world', [2, 3])
def test_halo_exchange_gloo(world):
    mp = pytest.importorskip('torch.multiprocessing')
    port = _free_port()
    mp.spawn(_spawn_cpu, args=(world, port), nprocs=world, join=True)


@pytest.mark.gpu
def test_decomposed_rk3_matches_single_gpu():
    torch = pytest.importorskip('torch')
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dist_worker.py'), 'gpu']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'DIST_GPU_OK' in out.stdout
