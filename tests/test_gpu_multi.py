"""Two (or more) real GPUs: the decomposed swarm with the peer-memory exchange (CUDA IPC inboxes, graph-replayed
refresh exchanges) and with NCCL send/recv against the undecomposed swarm. Skipped on a one-GPU box."""
import os, subprocess, sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize('transport,scenario', [('peer', 'none'), ('nccl', 'none'), ('peer', 'chase'), ('peer', 'net')])
def test_decomposed_swarm_on_real_gpus(transport, scenario):
    n = _n_gpus()
    if n < 2:
        pytest.skip('needs at least two GPUs')
    world = 4 if n >= 4 else 2
    port = 29640 + (0 if transport == 'peer' else 1) + {'none': 0, 'chase': 2, 'net': 4}[scenario]
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', str(port),
                        os.path.join(ROOT, 'tests', 'multi_gpu_worker.py'), transport, scenario],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert 'multi-gpu-ok' in r.stdout, r.stdout[-2000:]
