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
scenario = sys.argv[2] if len(sys.argv) > 2 else 'none'   # none | chase (one chasing predator, kill rule 3) | net (fish net)
world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
N, L, K = 120000, 3400.0, 130
if scenario == 'chase':
    d = base_dict('natPred', N, BC=0, size=L, Npred=1, pred_kill=3, pred_move=1, pred_time=0.02, time=50.0, pred_speed0=40.0,
                  kill_rate=400.0, N_confu=500)
elif scenario == 'net':
    d = base_dict('mulFisher', N, BC=0, size=L, Npred=200, pred_kill=1, pred_move=0, pred_time=0.02, time=50.0, pred_speed0=40.0)
else:
    d = base_dict('natPred', N, BC=0, size=L, Npred=0)
sp, npred = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
rng = np.random.default_rng(99)
st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=4.0 * np.ones(N))
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
slab = SlabSwarm(sp, N, rank, world, seed=7, device=local, stream=ts.cuda_stream, n_pred=npred)
slab.scatter_initial(st)
tr = PeerTransport(slab) if transport == 'peer' else DistTransport(rank, world)
step_distributed(slab, tr, 0, 40)
step_distributed(slab, tr, 40, K - 40)      # a second call: the exchange state carries over
torch.cuda.synchronize()
gid, own = slab.owned_state()
c = slab.counts()
stats = slab.swarm.stats()
gathered = [None] * world
pred_here = slab.swarm.get_predators()
dist.gather_object((gid, own, c, stats['cell_builds'], pred_here), gathered if rank == 0 else None, dst=0)
if rank == 0:
    whole = Swarm(sp, N, npred, seed=7, device=local)
    whole.set_state(**st)
    whole.step(0, K)
    ws = whole.get_state()
    wpred = whole.get_predators()
    whole.close()
    seen = np.zeros(N, dtype=np.int32)
    got = {k: np.zeros(N, dtype=v.dtype) for k, v in gathered[0][1].items()}
    for g_, o_, c_, builds, pr_ in gathered:
        assert np.allclose(pr_[..., :2], wpred[..., :2], atol=2e-3), 'every rank holds the whole swarm\'s predator'
        assert np.array_equal(pr_, gathered[0][4]), 'the replicated predator is bit-identical on all ranks'
        assert c_['overflow_events'] == 0, c_
        assert builds < K // 2, builds      # most exchanges were refreshes
        seen[g_] += 1
        for k, v in o_.items():
            got[k][g_] = v
    al = ws['alive'] == 1
    assert np.array_equal(seen == 1, al) and seen.max() <= 1, 'every living agent is owned by exactly one rank, the same agents died'
    if npred:
        assert (~al).sum() > 0, 'the scenario should kill somebody'
    assert np.array_equal(got['steps_till_burst'][al], ws['steps_till_burst'][al])
    assert np.array_equal(got['bin_step'][al], ws['bin_step'][al])
    dx = np.abs(((got['x'] - ws['x'] + L / 2) % L) - L / 2)
    dy = np.abs(((got['y'] - ws['y'] + L / 2) % L) - L / 2)
    close = ((dx < 1e-3) & (dy < 1e-3))[al]
    assert close.mean() > 0.995, close.mean()
    print('multi-gpu-ok world=%d transport=%s scenario=%s close=%.5f killed=%d migrated=%s' %
          (world, transport, scenario, close.mean(), int((~al).sum()), [int(x[2]['owned']) for x in gathered]))
slab.close()
dist.barrier()
dist.destroy_process_group()
