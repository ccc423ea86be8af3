#!/usr/bin/env python
"""Where the predator phase of the decomposed swarm spends its time: C5 weak scaling (1,048,576 agents per GPU) with
  none     no predator
  walk     random-walk predator, no kill   (predator-phase rows kernel + exchange in front of it + move kernel)
  chase    chasing predator, no kill       (+ closest-prey pass and all-gather)
  chase3   chasing predator, kill rule 3   (+ classification pass, kill all-gather)
usage: python -m torch.distributed.run --nproc-per-node W ... profiles/dom_pred_probe.py [variants...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch, torch.distributed as dist
import bench
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.domain import SlabSwarm, PeerTransport, step_distributed

world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
base = dict(pred_time=0.02, time=50.0, pred_speed0=40.0, kill_rate=400.0, N_confu=500)
variants = {'none': dict(Npred=0), 'walk': dict(base, Npred=1, pred_kill=0, pred_move=0), 'chase': dict(base, Npred=1, pred_kill=0, pred_move=1),
            'chase3': dict(base, Npred=1, pred_kill=3, pred_move=1)}
N = (1 << 20) * world
L = float(np.sqrt(N / bench.C5_RHO))
for name in (sys.argv[1:] or list(variants)):
    sp, npred = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, **variants[name]), path=PP.PATH_CELLLIST)
    slab = SlabSwarm(sp, N, rank, world, seed=bench.SEED, device=local, stream=ts.cuda_stream, n_pred=npred)
    slab.scatter_initial(bench.c5_state(N, L))
    tr = PeerTransport(slab)
    step_distributed(slab, tr, 0, 40)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    step_distributed(slab, tr, 40, 200)
    e1.record(ts)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    slab.close()
    if rank == 0:
        print(json.dumps({'variant': name, 'n_gpus': world, 'us_per_step': 1e3 * float(t.item()) / 200}), flush=True)
dist.barrier()
dist.destroy_process_group()
