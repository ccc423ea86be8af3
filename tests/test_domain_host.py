"""Host-side routing of the domain decomposition on the CPU: slab geometry, and the ring exchange of
record buffers through torch.distributed with the gloo backend (world sizes 2 and 3), i.e. the same
`DistTransport.exchange` code that moves device buffers over NCCL on the GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from sde_burst_coast_b200 import domain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_geometry():
    b = domain.slab_bounds(100.0, 4)
    assert b[0] == (0.0, 25.0) and b[-1][1] == 100.0
    x = np.array([0.0, 24.999, 25.0, 99.999, 50.0])
    assert list(domain.owner_of(x, 100.0, 4)) == [0, 0, 1, 3, 2]
    assert domain.neighbours(0, 4) == (3, 1) and domain.neighbours(3, 4) == (2, 0) and domain.neighbours(0, 2) == (1, 1)
    domain.check_geometry(400.0, 19.5, 8)
    with pytest.raises(ValueError):
        domain.check_geometry(100.0, 19.5, 8)    # 12.5-wide slabs < halo
    with pytest.raises(ValueError):
        domain.check_geometry(70.0, 19.5, 2)     # both neighbours are one peer: needs two halos


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from sde_burst_coast_b200.domain import DistTransport, neighbours
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
tr = DistTransport(rank, world)
n = 4096
for it in range(3):
    send_l = torch.full((n,), 10 * rank + 1 + 100 * it, dtype=torch.int32)
    send_r = torch.full((n,), 10 * rank + 2 + 100 * it, dtype=torch.int32)
    recv_l, recv_r = torch.zeros(n, dtype=torch.int32), torch.zeros(n, dtype=torch.int32)
    tr.exchange(send_l, send_r, recv_l, recv_r)
    left, right = neighbours(rank, world)
    assert int(recv_l[0]) == 10 * left + 2 + 100 * it and bool((recv_l == recv_l[0]).all()), (rank, 'from left', int(recv_l[0]))
    assert int(recv_r[0]) == 10 * right + 1 + 100 * it and bool((recv_r == recv_r[0]).all()), (rank, 'from right', int(recv_r[0]))
assert tr.min_int(7 + 3 * ((rank + 1) %% world)) == 7   # the rebuild period every rank can afford
dist.barrier()
dist.destroy_process_group()
sys.stdout.write('rank-' + str(rank) + '-ok\n'); sys.stdout.flush()
''' % ROOT


@pytest.mark.parametrize('world', [2, 3])
def test_ring_exchange_gloo(world, tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    port = 29600 + world
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count('-ok') == world, r.stdout


def test_world_one_is_a_self_exchange():
    import torch
    tr = domain.DistTransport(0, 1)
    a, b = torch.arange(8), torch.arange(8) + 100
    rl, rr = torch.zeros(8, dtype=a.dtype), torch.zeros(8, dtype=a.dtype)
    tr.exchange(a, b, rl, rr)
    assert rr.equal(a) and rl.equal(b)
    assert tr.min_int(5) == 5


def test_reference_arm_runs_on_rank_zero_only():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '1'], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''
