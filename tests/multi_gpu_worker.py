"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): a periodic swarm
decomposed into x-slabs with the peer-memory exchange against the same swarm on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch, torch.distributed as dist
from helpers import base_dict, f32
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.domain import SlabSwarm, PeerTransport, DistTransport, step_distributed
from sde_burst_coast_b200.swarm import Swarm

transport = sys.argv[1] if len(sys.argv) > 1 else 'peer'
world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
N, L, K = 120000, 3400.0, 130
d = base_dict('natPred', N, BC=0, size=L, Npred=0)
sp, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
rng = np.random.default_rng(99)
st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=4.0 * np.ones(N))
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
slab = SlabSwarm(sp, N, rank, world, seed=7, device=local, stream=ts.cuda_stream)
slab.scatter_initial(st)
tr = PeerTransport(slab) if transport == 'peer' else DistTransport(rank, world)
step_distributed(slab, tr, 0, 40)
step_distributed(slab, tr, 40, K - 40)      # a second call: the exchange state carries over
torch.cuda.synchronize()
gid, own = slab.owned_state()
c = slab.counts()
stats = slab.swarm.stats()
gathered = [None] * world
dist.gather_object((gid, own, c, stats['cell_builds']), gathered if rank == 0 else None, dst=0)
if rank == 0:
    whole = Swarm(sp, N, seed=7, device=local)
    whole.set_state(**st)
    whole.step(0, K)
    ws = whole.get_state()
    whole.close()
    seen = np.zeros(N, dtype=np.int32)
    got = {k: np.zeros(N, dtype=v.dtype) for k, v in gathered[0][1].items()}
    for g_, o_, c_, builds in gathered:
        assert c_['overflow_events'] == 0, c_
        assert builds < K // 2, builds      # most exchanges were refreshes
        seen[g_] += 1
        for k, v in o_.items():
            got[k][g_] = v
    assert np.all(seen == 1), 'every agent is owned by exactly one rank'
    assert np.array_equal(got['steps_till_burst'], ws['steps_till_burst'])
    assert np.array_equal(got['bin_step'], ws['bin_step'])
    dx = np.abs(((got['x'] - ws['x'] + L / 2) % L) - L / 2)
    dy = np.abs(((got['y'] - ws['y'] + L / 2) % L) - L / 2)
    close = (dx < 1e-3) & (dy < 1e-3)
    assert close.mean() > 0.995, close.mean()
    print('multi-gpu-ok world=%d transport=%s close=%.5f migrated=%s' % (world, transport, close.mean(), [int(x[2]['owned']) for x in gathered]))
slab.close()
dist.barrier()
dist.destroy_process_group()
